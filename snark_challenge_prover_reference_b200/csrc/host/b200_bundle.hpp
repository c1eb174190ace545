// B200-backed implementation of the reference's `B::` bundle API.
//
// The reference declares two classes of static functions over opaque types, mnt4753_libsnark and mnt6753_libsnark
// (libsnark/prover_reference_include/prover_reference_functions.hpp:5-83 and 84-162), and its prover driver
// (cuda_prover_piecewise.cu:18-98) is templated over such a bundle. This header provides the same two names with the
// same nested type names and the same static signatures, so that driver compiles unchanged against it; underneath,
// every vector lives in B200 HBM and every operation is a call into the C ABI of include/b200_groth16.h.
//
// Both curves share one class template; `mnt4753_libsnark` / `mnt6753_libsnark` are its two instantiations.
#pragma once
#include <cstddef>
#include <memory>
#include <vector>

namespace b200_host {
struct device_mem;   // owning device allocation
struct params_box;   // owning b200_params handle
struct domain_box;   // owning b200_domain handle
}  // namespace b200_host

template <int CURVE>
class b200_groth16_bundle {
public:
  // a view (shared storage + element offset) like the reference's {shared_ptr<vector>, offset}
  // (prover_reference_functions.cpp:135-145)
  struct vector_Fr {
    std::shared_ptr<b200_host::device_mem> data;
    size_t offset;
  };
  struct vector_G1 {
    std::shared_ptr<b200_host::params_box> owner;
    const void *data;  // device pointer to affine wire-format points
    int query;         // which query of the key this is (0 A, 1 B1, 3 L, 4 H)
  };
  struct vector_G2 {
    std::shared_ptr<b200_host::params_box> owner;
    const void *data;
    int query;         // 2 = B2
  };
  struct field {
    unsigned char bytes[96];  // Fr, Montgomery form
  };
  // Group elements live in host memory. One returned by multiexp_* may still be PENDING: the GPU work is enqueued, the
  // bytes arrive when the element is first read (resolve()), so the reference's driver - which issues its five
  // multiexps back to back (cuda_prover_piecewise.cu:71-81) - overlaps them without being changed.
  struct G1 {
    unsigned char bytes[3 * 96];  // projective (X:Y:Z) over Fq
    void *pending = nullptr;      // b200_msm_pending*
    void resolve();
    ~G1() { resolve(); }
  };
  struct G2 {
    unsigned char bytes[3 * 96 * (CURVE == 0 ? 2 : 3)];  // projective over Fq2 / Fq3
    void *pending = nullptr;
    void resolve();
    ~G2() { resolve(); }
  };
  struct evaluation_domain {
    std::shared_ptr<b200_host::domain_box> box;
  };
  class groth16_params {
  public:
    size_t d, m;
    std::shared_ptr<b200_host::params_box> box;
  };
  class groth16_input {
  public:
    std::shared_ptr<b200_host::device_mem> w, ca, cb, cc;
    field r;
  };

  static void init_public_params();

  static void print_G1(G1 *a);
  static void print_G2(G2 *a);

  static evaluation_domain *get_evaluation_domain(size_t d);

  static G1 *G1_add(G1 *a, G1 *b);
  static G1 *G1_scale(field *a, G1 *b);

  static void vector_Fr_muleq(vector_Fr *a, vector_Fr *b, size_t size);
  static void vector_Fr_subeq(vector_Fr *a, vector_Fr *b, size_t size);
  static vector_Fr *vector_Fr_offset(vector_Fr *a, size_t offset);
  static void vector_Fr_copy_into(vector_Fr *src, vector_Fr *dst, size_t length);
  static vector_Fr *vector_Fr_zeros(size_t length);

  static void domain_iFFT(evaluation_domain *domain, vector_Fr *a);
  static void domain_cosetFFT(evaluation_domain *domain, vector_Fr *a);
  static void domain_icosetFFT(evaluation_domain *domain, vector_Fr *a);
  static void domain_divide_by_Z_on_coset(evaluation_domain *domain, vector_Fr *a);
  static size_t domain_get_m(evaluation_domain *domain);

  static G1 *multiexp_G1(vector_Fr *scalar_start, vector_G1 *g_start, size_t length);
  static G2 *multiexp_G2(vector_Fr *scalar_start, vector_G2 *g_start, size_t length);

  static groth16_input *read_input(const char *path, groth16_params *params);

  static vector_Fr *input_w(groth16_input *input);
  static vector_Fr *input_ca(groth16_input *input);
  static vector_Fr *input_cb(groth16_input *input);
  static vector_Fr *input_cc(groth16_input *input);
  static field *input_r(groth16_input *input);

  static groth16_params *read_params(const char *path);

  static size_t params_d(groth16_params *params);
  static size_t params_m(groth16_params *params);
  static vector_G1 *params_A(groth16_params *params);
  static vector_G1 *params_B1(groth16_params *params);
  static vector_G1 *params_L(groth16_params *params);
  static vector_G1 *params_H(groth16_params *params);
  static vector_G2 *params_B2(groth16_params *params);

  static void delete_G1(G1 *a);
  static void delete_G2(G1 *a);  // the reference's signature takes a G1* (hpp:73); kept for source compatibility
  static void delete_G2(G2 *a);
  static void delete_vector_Fr(vector_Fr *a);
  static void delete_vector_G1(vector_G1 *a);
  static void delete_vector_G2(vector_G2 *a);
  static void delete_groth16_input(groth16_input *a);
  static void delete_groth16_params(groth16_params *a);
  static void delete_evaluation_domain(evaluation_domain *a);

  static void groth16_output_write(G1 *A, G2 *B, G1 *C, const char *output_path);
};

typedef b200_groth16_bundle<0> mnt4753_libsnark;
typedef b200_groth16_bundle<1> mnt6753_libsnark;
