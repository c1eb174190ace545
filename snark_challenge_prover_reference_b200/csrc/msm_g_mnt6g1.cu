// Mnt6G1 instantiation of the group-dependent MSM kernels (see msm_group.cuh).
#include "msm_group.cuh"
namespace b200 {
template int msm_run<Mnt6G1>(const void *, const void *, size_t, void *);
int msm_run_mnt6g1(const void *s, const void *p, size_t n, void *out) { return msm_run<Mnt6G1>(s, p, n, out); }
}  // namespace b200
