# Python model of the even/odd-lane interleaved Montgomery multiplication (32-bit limbs, carry flag), to validate the
# algorithm before writing it in PTX.
import random
L = 2**252 + 27742317777372353535851937790883648493
MASK = 0xffffffff
l = [(L >> (32*i)) & MASK for i in range(8)]
M0 = (-pow(L, -1, 2**32)) % 2**32
R = 2**256

class CC:
    def __init__(s): s.cf = 0
    def add_cc(s, a, b): t = a + b; s.cf = t >> 32; return t & MASK
    def addc_cc(s, a, b): t = a + b + s.cf; s.cf = t >> 32; return t & MASK
    def addc(s, a, b): t = a + b + s.cf; return t & MASK   # cf unchanged (no .cc) but we don't rely on it afterwards
    def mad_lo_cc(s, a, b, c): t = ((a*b) & MASK) + c; s.cf = t >> 32; return t & MASK
    def madc_lo_cc(s, a, b, c): t = ((a*b) & MASK) + c + s.cf; s.cf = t >> 32; return t & MASK
    def madc_hi_cc(s, a, b, c): t = ((a*b) >> 32) + c + s.cf; s.cf = t >> 32; return t & MASK
    def madc_hi(s, a, b, c): t = ((a*b) >> 32) + c + s.cf; assert t >> 32 == 0, "overflow in final madc.hi"; return t & MASK

def cmad4(cc, acc, x, y, top):
    """acc[0..7] lanes += x[k]*y (k=0..3) as one carry chain; carry-out added to top[7]"""
    acc[0] = cc.mad_lo_cc(x[0], y, acc[0]); acc[1] = cc.madc_hi_cc(x[0], y, acc[1])
    for k in range(1, 4):
        acc[2*k] = cc.madc_lo_cc(x[k], y, acc[2*k]); acc[2*k+1] = cc.madc_hi_cc(x[k], y, acc[2*k+1])
    t = top[7] + cc.cf; assert t >> 32 == 0; top[7] = t

def redc_step(cc, even, odd):
    mi = (even[0] * M0) & MASK
    # odd lanes (words 1,3,5,7) += l1*mi, l3*mi, (l5 = 0), l7*mi
    odd[0] = cc.mad_lo_cc(l[1], mi, odd[0]); odd[1] = cc.madc_hi_cc(l[1], mi, odd[1])
    odd[2] = cc.madc_lo_cc(l[3], mi, odd[2]); odd[3] = cc.madc_hi_cc(l[3], mi, odd[3])
    odd[4] = cc.addc_cc(odd[4], 0); odd[5] = cc.addc_cc(odd[5], 0)
    odd[6] = cc.madc_lo_cc(l[7], mi, odd[6]); odd[7] = cc.madc_hi(l[7], mi, odd[7])
    # even lanes (words 0,2,4,6) += l0*mi, l2*mi, (l4 = l6 = 0); carry into word 8 = odd[7]
    even[0] = cc.mad_lo_cc(l[0], mi, even[0]); even[1] = cc.madc_hi_cc(l[0], mi, even[1])
    even[2] = cc.madc_lo_cc(l[2], mi, even[2]); even[3] = cc.madc_hi_cc(l[2], mi, even[3])
    for k in range(4, 8): even[k] = cc.addc_cc(even[k], 0)
    t = odd[7] + cc.cf; assert t >> 32 == 0; odd[7] = t
    assert even[0] == 0

def montmul(a, b):
    A = [(a >> (32*i)) & MASK for i in range(8)]; B = [(b >> (32*i)) & MASK for i in range(8)]
    cc = CC()
    even = [0]*8; odd = [0]*8
    # first row: plain products
    for k in range(4):
        p = A[2*k] * B[0]; even[2*k] = p & MASK; even[2*k+1] = p >> 32
        p = A[2*k+1] * B[0]; odd[2*k] = p & MASK; odd[2*k+1] = p >> 32
    redc_step(cc, even, odd)
    for i in range(1, 8):
        even, odd = odd, even      # role swap = division by 2^32
        bi = B[i]
        # merge the dead lane's high word, then rebuild the odd lanes shifted right by two words
        even[0] = cc.add_cc(even[0], odd[1])
        odd[0] = cc.madc_lo_cc(A[1], bi, odd[2]); odd[1] = cc.madc_hi_cc(A[1], bi, odd[3])
        odd[2] = cc.madc_lo_cc(A[3], bi, odd[4]); odd[3] = cc.madc_hi_cc(A[3], bi, odd[5])
        odd[4] = cc.madc_lo_cc(A[5], bi, odd[6]); odd[5] = cc.madc_hi_cc(A[5], bi, odd[7])
        odd[6] = cc.madc_lo_cc(A[7], bi, 0); odd[7] = cc.madc_hi(A[7], bi, 0)
        cmad4(cc, even, [A[0], A[2], A[4], A[6]], bi, odd)
        redc_step(cc, even, odd)
    # result = S / 2^32: words 1..8: odd[j] (word j+1... ) merge
    # now 'even' holds lanes at words 0,2,4,6 (even[0] == 0), 'odd' lanes at words 1,3,5,7
    res = [0]*8
    res[0] = cc.add_cc(odd[0], even[1])
    for j in range(1, 7): res[j] = cc.addc_cc(odd[j], even[j+1])
    t = odd[7] + cc.cf; assert t >> 32 == 0; res[7] = t
    r = sum(res[j] << (32*j) for j in range(8))
    if r >= L: r -= L
    assert r < L
    return r

random.seed(1)
Rinv = pow(R, -1, L)
for t in range(20000):
    a = random.randrange(L); b = random.randrange(L)
    if t < 10: a = L - 1 - t; b = L - 1 - (t % 3)
    assert montmul(a, b) == a * b * Rinv % L, t
print("ok")
