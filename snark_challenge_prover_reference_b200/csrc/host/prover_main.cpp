// Host driver with the reference's command line:  <MNT4753|MNT6753> compute <params> <input> <output>
// It drives the prover exclusively through the `B::` bundle, in the order of the reference's
// cuda_prover_piecewise.cu:18-98 (witness map, five multi-exponentiations, C = H + L + r*B1, write A|B|C), and prints
// the same style of phase timings as libsnark/main.cpp:201-270 ("Total time from input to output" starts after the
// parameters are loaded). The reference's own unmodified driver also compiles against prover_reference_functions.hpp
// (oracle/build_ref.sh builds it as oracle/_ref/piecewise_b200); this file exists so that the repo is self-contained.
#include <chrono>
#include <cstdio>
#include <string>

#include "b200_bundle.hpp"

typedef std::chrono::steady_clock clk;
static double since_ms(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

template <typename B>
static typename B::vector_Fr *witness_map(size_t d, typename B::vector_Fr *ca, typename B::vector_Fr *cb,
                                          typename B::vector_Fr *cc) {
  typename B::evaluation_domain *dom = B::get_evaluation_domain(d + 1);
  const size_t m = B::domain_get_m(dom);
  B::domain_iFFT(dom, ca);
  B::domain_iFFT(dom, cb);
  B::domain_cosetFFT(dom, ca);
  B::domain_cosetFFT(dom, cb);
  B::vector_Fr_muleq(ca, cb, m);  // ca <- ca .* cb on the coset
  B::domain_iFFT(dom, cc);
  B::domain_cosetFFT(dom, cc);
  B::vector_Fr_subeq(ca, cc, m);
  B::domain_divide_by_Z_on_coset(dom, ca);
  B::domain_icosetFFT(dom, ca);
  typename B::vector_Fr *h = B::vector_Fr_zeros(m + 1);
  B::vector_Fr_copy_into(ca, h, m);
  B::delete_evaluation_domain(dom);
  return h;
}

template <typename B>
static int prove(const char *params_path, const char *input_path, const char *output_path) {
  B::init_public_params();
  clk::time_point t0 = clk::now();
  typename B::groth16_params *params = B::read_params(params_path);
  printf("load params: %.0f ms\n", since_ms(t0));
  clk::time_point t_main = clk::now();
  typename B::groth16_input *input = B::read_input(input_path, params);
  printf("load inputs: %.0f ms\n", since_ms(t_main));
  const size_t d = B::params_d(params), m = B::params_m(params);

  clk::time_point t = clk::now();
  typename B::vector_Fr *ca = B::input_ca(input), *cb = B::input_cb(input), *cc = B::input_cc(input);
  typename B::vector_Fr *h = witness_map<B>(d, ca, cb, cc);
  printf("compute H: %.1f ms\n", since_ms(t));

  typename B::vector_Fr *w = B::input_w(input);
  typename B::vector_Fr *w2 = B::vector_Fr_offset(w, 2);  // primary_input_size + 1
  typename B::vector_G1 *qA = B::params_A(params), *qB1 = B::params_B1(params), *qL = B::params_L(params),
                        *qH = B::params_H(params);
  typename B::vector_G2 *qB2 = B::params_B2(params);
  t = clk::now();
  typename B::G1 *At = B::multiexp_G1(w, qA, m + 1);
  printf("A G1 multiexp issued: %.1f ms\n", since_ms(t));
  t = clk::now();
  typename B::G1 *Bt1 = B::multiexp_G1(w, qB1, m + 1);
  printf("B G1 multiexp issued: %.1f ms\n", since_ms(t));
  t = clk::now();
  typename B::G2 *Bt2 = B::multiexp_G2(w, qB2, m + 1);
  printf("B G2 multiexp issued: %.1f ms\n", since_ms(t));
  t = clk::now();
  typename B::G1 *Ht = B::multiexp_G1(h, qH, d);
  printf("H G1 multiexp issued: %.1f ms\n", since_ms(t));
  t = clk::now();
  typename B::G1 *Lt = B::multiexp_G1(w2, qL, m - 1);
  printf("L G1 multiexp issued: %.1f ms\n", since_ms(t));

  typename B::field *r = B::input_r(input);
  typename B::G1 *rB = B::G1_scale(r, Bt1);
  typename B::G1 *LrB = B::G1_add(Lt, rB);
  typename B::G1 *C = B::G1_add(Ht, LrB);
  // (the five multiexps above are asynchronous: their results are first read by G1_scale / G1_add / the writer)
  B::groth16_output_write(At, Bt2, C, output_path);
  printf("Total time from input to output: : %.0f ms\n", since_ms(t_main));

  B::delete_G1(At); B::delete_G1(Bt1); B::delete_G2(Bt2); B::delete_G1(Ht); B::delete_G1(Lt);
  B::delete_G1(rB); B::delete_G1(LrB); B::delete_G1(C);
  delete r;
  B::delete_vector_Fr(ca); B::delete_vector_Fr(cb); B::delete_vector_Fr(cc); B::delete_vector_Fr(h);
  B::delete_vector_Fr(w); B::delete_vector_Fr(w2);
  B::delete_vector_G1(qA); B::delete_vector_G1(qB1); B::delete_vector_G1(qL); B::delete_vector_G1(qH);
  B::delete_vector_G2(qB2);
  B::delete_groth16_input(input);
  B::delete_groth16_params(params);
  return 0;
}

int main(int argc, char **argv) {
  setbuf(stdout, NULL);
  if (argc != 6 || std::string(argv[2]) != "compute") {
    fprintf(stderr, "usage: %s <MNT4753|MNT6753> compute <params> <input> <output>\n", argv[0]);
    return 2;
  }
  const std::string curve(argv[1]);
  if (curve == "MNT4753") return prove<mnt4753_libsnark>(argv[3], argv[4], argv[5]);
  if (curve == "MNT6753") return prove<mnt6753_libsnark>(argv[3], argv[4], argv[5]);
  fprintf(stderr, "unknown curve %s\n", argv[1]);
  return 2;
}
