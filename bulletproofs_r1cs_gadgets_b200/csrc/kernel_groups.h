// Kernel instantiation groups: each k_*.cu explicitly instantiates launch<K> (and with it the
// __global__ run_kernel<K>) for one group so the heavy kernels compile in parallel; every other
// translation unit sees them as extern templates.
#pragma once
#include "kernels.h"
#include "devrt.h"

#define KGROUP_MSM(X) X(KMsmAccumulate) X(KMsmFinish) X(KMsmWindowSum)
#define KGROUP_FOLD(X) X(KFoldGens) X(KFoldTable)
#define KGROUP_TABLE(X) X(KTableBuild) X(KMsmTable) X(KMsmTableFinish) X(KVerifyCheck) X(KMsmTableReduce) X(KVerifyProofPoint) X(KSumPointsStrided) X(KVerifyCombinedCheck)
#define KGROUP_POINTS(X) X(KCommit) X(KGensFromUniform) X(KPcBases) X(KPcTable) X(KEncodePoints) X(KVerifyDecompress)
#define KGROUP_TRANSCRIPT(X) X(KTsStart) X(KRngDraw) X(KTsPhase2) X(KTsPhase3) X(KTsPhase4) X(KTsIpaRound) X(KSelfTest) X(KTsVerify) X(KBatchSeed) X(KVerifyRho)
#define KGROUP_SCALAR(X) X(KLoadScalars) X(KRecode) X(KPowers) X(KFillScalar) X(KFlatten) X(KFlattenParts) X(KFlattenSum) X(KPolyT) X(KSumPartials) X(KPolyEval) \
  X(KProverScalars) X(KIpaDots) X(KRecodeIpa) X(KFoldAB) X(KStoreAB) X(KWitnessTape) X(KVerifyS) X(KVerifyDelta) X(KVerifyGH) X(KVerifyScalars) X(KVerifyCombineRows) X(KCheckFixedCommitments) X(KIpaUTable) X(KRecodeUnfolded) X(KRecodeFoldTable)

#define KGROUP_SORTED(X) X(KShiftTableBuild) X(KShiftTableOne) X(KMergeGens) X(KRecode13) X(KRecodeUnfolded13) X(KSortBucketsSerial) X(KBucketAccumulate) X(KBucketAccumulateRel) X(KBucketReduce) X(KBucketReduceW) X(KBucketFinishA) X(KBucketFinishB) X(KSumPointsEncode)

#define KDECL_EXTERN(K) extern template int launch<K>(long, dev_stream, const K &);
#define KDEFINE(K) template int launch<K>(long, dev_stream, const K &);

#ifndef KGROUP_DEFINING
KGROUP_MSM(KDECL_EXTERN)
KGROUP_FOLD(KDECL_EXTERN)
KGROUP_TABLE(KDECL_EXTERN)
KGROUP_SORTED(KDECL_EXTERN)
KGROUP_POINTS(KDECL_EXTERN)
KGROUP_TRANSCRIPT(KDECL_EXTERN)
KGROUP_SCALAR(KDECL_EXTERN)
#endif

// block-cooperative counting sort of digit items by bucket (k_sorted.cu)
// tmp (optional): scratch of ninst * items_stride words for the two-pass form; gen_slots: generator slots of the shift table
int launch_sort_buckets(const RowMap &rmap, const int8_t *dig, long dig_inst_stride, long rows, long ninst, uint32_t *items, long items_stride,
                        uint32_t *boff, uint32_t *soff, uint32_t *tmp, size_t tmp_bytes, long gen_slots, dev_stream s);
