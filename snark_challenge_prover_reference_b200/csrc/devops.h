// Host-callable entry points of devops.cu (internal; the public surface is include/b200_groth16.h).
#pragma once
#include <stddef.h>
#include <stdint.h>
namespace b200 {
int dev_fp_op(int tag, int op, const void *a, const void *b, void *r, size_t n);
int dev_fqe_op(int curve, int op, const void *a, const void *b, void *r, size_t n);
int dev_group_op(int curve, int group, int op, const void *p, const void *q, void *r, size_t n);
int gen_points(int curve, int group, void *out, size_t n, uint64_t first);
int batch_exp(int curve, int group, const void *h_base, const void *d_scalars, size_t n, void *d_out, int window, double *ms3);
int imad_peak(double *mac32_per_s2, double *ms2);
}  // namespace b200
