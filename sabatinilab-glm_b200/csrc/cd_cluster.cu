// C-ABI entry of the cluster coordinate-descent kernel (kernel: cd_cluster.cuh; instantiations per
// group size: cd_cluster_m{1,2,4,8}.cu).
#include <stdlib.h>

#include "cd_cluster.cuh"

using namespace sglm;

extern "C" int sglm_enet_cd_cluster_supported(int32_t group_size, int32_t cluster_size) {
    const int M = group_size, K = cluster_size;
    if (M == 8) return (K == 2 || K == 4 || K == 8) ? 1 : 0;
    return ((M == 1 || M == 2 || M == 4) && (K == 1 || K == 2 || K == 4 || K == 8)) ? 1 : 0;
}

typedef CUresult (*PFN_encodeTiled_cd)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                       const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                       CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                       CUtensorMapFloatOOBfill);

extern "C" size_t sglm_enet_cd_cluster_tmap_bytes(void) { return sizeof(CUtensorMap); }

// One TMA descriptor per problem: Q as a C x C fp64 matrix with row pitch ldq, boxes of 32 x 32, zero fill
// outside.  Encoded on the host into `out` (n_prob * sglm_enet_cd_cluster_tmap_bytes()); the caller copies
// it to device memory (64-byte aligned) and passes that as prob_tmap.
extern "C" int sglm_enet_cd_cluster_encode_tmaps(const uint64_t *Q_dev_ptrs, int32_t n_prob, int32_t C, int64_t ldq,
                                                 void *out) {
    SGLM_CHECK_ARG(Q_dev_ptrs && out && n_prob >= 0 && C > 0 && ldq >= C, SGLM_E_INVALID_ARG, "cd tmaps: bad argument");
    SGLM_CHECK_ARG((ldq & 1) == 0, SGLM_E_ALIGN, "cd tmaps: ldq must be even (16-byte rows)");
    static PFN_encodeTiled_cd enc = nullptr;
    if (!enc) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            enc = (PFN_encodeTiled_cd)p;
    }
    SGLM_CHECK_ARG(enc != nullptr, SGLM_E_CUDA, "cd tmaps: cuTensorMapEncodeTiled not available");
    CUtensorMap *tm = (CUtensorMap *)out;
    for (int i = 0; i < n_prob; ++i) {
        SGLM_CHECK_ARG((Q_dev_ptrs[i] & 15) == 0, SGLM_E_ALIGN, "cd tmaps: Q must be 16-byte aligned");
        const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)C};
        const cuuint64_t gstride[1] = {(cuuint64_t)ldq * 8};
        const cuuint32_t box[2] = {32, 32};
        const cuuint32_t estr[2] = {1, 1};
        CUtensorMap t;
        CUresult r = enc(&t, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void *)Q_dev_ptrs[i], gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        SGLM_CHECK_ARG(r == CUDA_SUCCESS, SGLM_E_CUDA, "cd tmaps: cuTensorMapEncodeTiled failed (%d)", (int)r);
        memcpy(tm + i, &t, sizeof(CUtensorMap));
    }
    return SGLM_OK;
}

// Shared memory one CTA of the (group_size, cluster_size) shape needs for a design of C columns (0 = shape not
// compiled in): the host picks a larger cluster (narrower column slice per CTA) or the per-model kernel when
// this exceeds the 227 KB a CTA can have.
extern "C" size_t sglm_enet_cd_cluster_smem_bytes(int32_t group_size, int32_t cluster_size, int32_t C) {
    if (!sglm_enet_cd_cluster_supported(group_size, cluster_size) || C <= 0) return 0;
    return cdc::make_layout(group_size, cluster_size, 8, C).total;
}

extern "C" int sglm_enet_cd_cluster_f64(const double *const *prob_Q, const double *const *prob_q,
                                        const double *const *prob_diag, const double *prob_yy, int64_t ldq,
                                        int32_t C, const int32_t *prob_of_group, const int32_t *model_of_slot,
                                        const double *l1_reg, const double *l2_reg, const double *tol,
                                        const int32_t *max_iter, int32_t n_groups, int32_t group_size,
                                        int32_t cluster_size, int32_t warm_start, int32_t do_screening, double *W,
                                        int64_t ldw, double *info, const void *prob_tmap, double *group_stats,
                                        void *stream) {
    SGLM_CHECK_ARG(C > 0 && n_groups >= 0 && ldq >= C && ldw >= C, SGLM_E_SHAPE, "enet_cd_cluster: bad shape");
    if (n_groups == 0) return SGLM_OK;
    SGLM_CHECK_ARG(prob_Q && prob_q && prob_diag && prob_yy && prob_of_group && model_of_slot && l1_reg && l2_reg &&
                       tol && max_iter && W && info && prob_tmap,
                   SGLM_E_INVALID_ARG, "enet_cd_cluster: null pointer");
    SGLM_CHECK_ARG((ldq & 1) == 0, SGLM_E_ALIGN, "enet_cd_cluster: ldq must be even (16-byte rows)");
    SGLM_CHECK_ARG(sglm_enet_cd_cluster_supported(group_size, cluster_size), SGLM_E_UNSUPPORTED,
                   "enet_cd_cluster: unsupported (group_size=%d, cluster_size=%d)", group_size, cluster_size);
    SGLM_CHECK_ARG((C + 31) / 32 >= cluster_size, SGLM_E_UNSUPPORTED,
                   "enet_cd_cluster: C=%d too small for a cluster of %d", C, cluster_size);
    cdc::Args a{prob_Q, prob_q, prob_diag, prob_yy, (long long)ldq, C, prob_of_group, model_of_slot, l1_reg, l2_reg,
                tol, max_iter, n_groups, warm_start, do_screening, W, (long long)ldw, info,
                (const CUtensorMap *)prob_tmap, (cudaStream_t)stream, 0, group_stats};
    if (const char *v = tuning_env("SGLM_CDC_VARIANT")) a.variant = atoi(v);      // tuning switch
    if (group_size == 1) return cdc::launch_m1(a, cluster_size);
    if (group_size == 2) return cdc::launch_m2(a, cluster_size);
    if (group_size == 8) return cdc::launch_m8(a, cluster_size);
    return cdc::launch_m4(a, cluster_size);
}
