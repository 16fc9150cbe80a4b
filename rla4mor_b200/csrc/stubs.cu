// stubs.cu -- entry points declared in include/rla_b200.h whose kernels are not
// written yet.  They fail loudly (RLA_ERR_UNSUPPORTED); they never fall back.
#include "common.cuh"
#define RLA_STUB(name) return ::rla::fail(RLA_ERR_UNSUPPORTED, name ": not implemented yet")

extern "C" {
int rla_embed_apply_rng_f32(uint64_t, int, float, int64_t, int64_t, int64_t, int64_t, const float *, int64_t, int64_t,
                            float *, int64_t, int, void *, size_t, void *) { RLA_STUB("rla_embed_apply_rng_f32"); }
int rla_gemm_nn_f64(const double *, int64_t, int64_t, int64_t, const double *, int64_t, int64_t, double *, int64_t, void *) { RLA_STUB("rla_gemm_nn_f64"); }
int rla_spmm_csr_f64(const int64_t *, const int32_t *, const double *, int64_t, int64_t, const double *, int64_t, int64_t,
                     double *, int64_t, void *) { RLA_STUB("rla_spmm_csr_f64"); }
int rla_gram_schmidt_f64(double *, int64_t, int64_t, int64_t, int64_t, double *, int32_t *, double, double, double, void *) { RLA_STUB("rla_gram_schmidt_f64"); }
int rla_svd_jacobi_f64(double *, int64_t, int64_t, int64_t, double *, double *, int, void *) { RLA_STUB("rla_svd_jacobi_f64"); }
int rla_residual_norm_f64(const double *, int64_t, int64_t, int64_t, const double *, const double *, const double *, int64_t,
                          const double *, double *, void *) { RLA_STUB("rla_residual_norm_f64"); }
}
