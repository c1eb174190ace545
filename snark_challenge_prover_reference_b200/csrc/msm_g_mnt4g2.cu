// Mnt4G2 instantiation of the group-dependent MSM kernels (see msm_group.cuh).
#include "msm_group.cuh"
namespace b200 {
int msm_run_mnt4g2(const void *s, const void *p, size_t n, void *out) { return msm_run<Mnt4G2>(s, p, n, out); }
int msm_run_deferred_mnt4g2(const void *s, const void *p, size_t n, void *out, std::function<void()> &tail) {
  return msm_run_deferred<Mnt4G2>(s, p, n, out, tail);
}
}  // namespace b200
