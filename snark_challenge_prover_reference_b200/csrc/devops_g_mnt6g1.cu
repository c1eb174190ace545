// Mnt6G1 instantiation of the group test/bench kernels (see devops_group.cuh).
#include "devops_group.cuh"
namespace b200 {
int group_op_mnt6g1(int op, const void *p, const void *q, void *r, size_t n) { return group_op_t<Mnt6G1>(op, p, q, r, n); }
int gen_points_mnt6g1(void *out, size_t n, uint64_t first) { return gen_points_t<Mnt6G1>(out, n, first); }
int batch_exp_mnt6g1(const void *g, const void *s, size_t n, void *out, int window, double *ms3) {
  return batch_exp_t<Mnt6G1>(g, s, n, out, window, ms3);
}
}  // namespace b200
