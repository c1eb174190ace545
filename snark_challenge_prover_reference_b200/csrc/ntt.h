// Host-callable entry points of ntt.cu (internal; the public surface is include/b200_groth16.h).
#pragma once
#include <stddef.h>
#include <cuda_runtime.h>
namespace b200 {
struct Domain;
int domain_create(int curve, size_t m, Domain **out);
void domain_destroy(Domain *d);
size_t domain_size(const Domain *d);
// test hook: first `count` entries of table 0 omega^i, 1 omega^-i, 2 g^i, 3 g^-i/m, 4 {1/m, 1/Z(g)}
int domain_table(const Domain *d, int which, void *h_out, size_t count);
// kind: 0 FFT, 1 iFFT, 2 cosetFFT, 3 icosetFFT
int domain_transform(Domain *d, void *d_a, int kind);
int domain_divide_by_z(Domain *d, void *d_a);
int fr_muleq(int curve, void *d_a, const void *d_b, size_t n);
int fr_subeq(int curve, void *d_a, const void *d_b, size_t n);
int fr_fill_pseudo_random(void *d_out, size_t n_elements, uint64_t seed);
int compute_h(Domain *d, void *d_ca, void *d_cb, void *d_cc, void *d_out);
// stream used by the transforms / point-wise kernels issued from the calling host thread (0 = legacy default)
void ntt_set_stream(cudaStream_t st);
}  // namespace b200
