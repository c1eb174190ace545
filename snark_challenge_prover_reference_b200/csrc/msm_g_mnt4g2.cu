// Mnt4G2 instantiation of the group-dependent MSM kernels (see msm_group.cuh).
#include "msm_group.cuh"
namespace b200 {
int msm_run_mnt4g2(const void *s, const void *p, size_t n, void *out) { return msm_run<Mnt4G2>(s, p, n, out); }
int msm_run_deferred_mnt4g2(const void *s, const void *p, size_t n, void *out, MsmTail &tail) {
  return msm_run_deferred<Mnt4G2>(s, p, n, out, tail);
}
int msm_precompute_mnt4g2(const void *p, size_t n, MsmPlan &plan, DevBuf &table) { return msm_precompute<Mnt4G2>(p, n, plan, table); }
int msm_run_table_deferred_mnt4g2(const void *s, const void *t, size_t n, const MsmPlan &plan, void *out,
                                  MsmTail &tail, MsmShare share, const MsmDedup *dedup) {
  return msm_run_table_deferred<Mnt4G2>(s, t, n, plan, out, tail, share, dedup);
}
}  // namespace b200
