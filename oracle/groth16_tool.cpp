// Oracle infrastructure (NOT product code; links the UNMODIFIED reference): a verification harness for complete
// Groth16 proofs, built from the reference's own generator, serialisers and pairing-based verifier.
//
//   groth16_tool gen    <MNT4753|MNT6753> <log2(d+1)> <params> <input> <extras> <vk>
//   groth16_tool verify <MNT4753|MNT6753> <vk> <input> <proof>
//
// `gen` does what libsnark/generate_parameters.cpp:23-123 does - the same example circuit, the same generator call, the
// same params / input files through the same write_* functions (libsnark/serialization.hpp) - and ADDITIONALLY writes
//   <extras>: alpha_g1 | beta_g1 | delta_g1 (G1) | beta_g2 | delta_g2 (G2), affine wire format - the proving-key elements
//             the challenge's parameter file leaves out (r1cs_gg_ppzksnark.hpp:72-76), which a complete proof needs;
//   <vk>    : the verification key in libsnark's own text format (the file main.cpp:331-335 would read).
// (generate_parameters.cpp only writes those under `const bool debug = false`, :21,110-120, so the unmodified generator
// binary cannot produce them.)
// `verify` reads a COMPLETE proof (A in G1 | B in G2 | C in G1, wire format, i.e. after the alpha / beta / delta / r / s
// terms of r1cs_gg_ppzksnark.tcc:457-470 have been added - main.cpp:295-343 `debug` sketches exactly this), takes the
// primary input w[1] from the input file and runs r1cs_gg_ppzksnark_verifier_strong_IC (r1cs_gg_ppzksnark.tcc:595-614).
// Exit status 0 = the proof verifies.
#include <cassert>
#include <cstdio>
#include <fstream>
#include <string>

#include <libff/common/profiling.hpp>
#include <libff/common/rng.hpp>
#include <libff/common/utils.hpp>
#include <libsnark/serialization.hpp>
#include <libff/algebra/curves/mnt753/mnt4753/mnt4753_pp.hpp>
#include <libff/algebra/curves/mnt753/mnt6753/mnt6753_pp.hpp>
#include <omp.h>
#include <libff/algebra/scalar_multiplication/multiexp.hpp>
#include <libsnark/knowledge_commitment/kc_multiexp.hpp>
#include <libsnark/reductions/r1cs_to_qap/r1cs_to_qap.hpp>
#include <libsnark/relations/constraint_satisfaction_problems/r1cs/examples/r1cs_examples.hpp>
#include <libsnark/zk_proof_systems/ppzksnark/r1cs_gg_ppzksnark/r1cs_gg_ppzksnark.hpp>

using namespace libsnark;
using namespace libff;

template <typename ppT>
static int gen(int log2_d, const char *params_path, const char *input_path, const char *extras_path, const char *vk_path) {
  srand(time(NULL));
  ppT::init_public_params();
  libff::inhibit_profiling_info = true;
  libff::inhibit_profiling_counters = true;
  const size_t primary_input_size = 1;
  const size_t d_plus_1 = (size_t)1 << log2_d, d = d_plus_1 - 1;
  // generate_parameters.cpp:37-56
  r1cs_example<Fr<ppT>> example = generate_r1cs_example_with_field_input<Fr<ppT>>(d - 1, 1);
  r1cs_gg_ppzksnark_keypair<ppT> keypair = r1cs_gg_ppzksnark_generator<ppT>(example.constraint_system);
  r1cs_variable_assignment<Fr<ppT>> full = example.primary_input;
  full.insert(full.end(), example.auxiliary_input.begin(), example.auxiliary_input.end());
  std::vector<Fr<ppT>> ca(d_plus_1, Fr<ppT>::zero()), cb(d_plus_1, Fr<ppT>::zero()), cc(d_plus_1, Fr<ppT>::zero());
  const size_t nc = keypair.pk.constraint_system.num_constraints();
  for (size_t i = 0; i <= primary_input_size; ++i) ca[i + nc] = (i > 0 ? full[i - 1] : Fr<ppT>::one());
  for (size_t i = 0; i < nc; ++i) {
    ca[i] += keypair.pk.constraint_system.constraints[i].a.evaluate(full);
    cb[i] += keypair.pk.constraint_system.constraints[i].b.evaluate(full);
    cc[i] += keypair.pk.constraint_system.constraints[i].c.evaluate(full);
  }
  // generate_parameters.cpp:58-86
  const size_t m = keypair.pk.constraint_system.num_variables();
  FILE *params = fopen(params_path, "w");
  write_size_t(params, d);
  write_size_t(params, m);
  for (size_t i = 0; i <= m; ++i) write_g1<ppT>(params, keypair.pk.A_query[i]);
  for (size_t i = 0; i <= m; ++i) write_g1<ppT>(params, keypair.pk.B_query[i].h);
  for (size_t i = 0; i <= m; ++i) write_g2<ppT>(params, keypair.pk.B_query[i].g);
  for (size_t i = 0; i < m - 1; ++i) write_g1<ppT>(params, keypair.pk.L_query[i]);
  for (size_t i = 0; i < d; ++i) write_g1<ppT>(params, keypair.pk.H_query[i]);
  fclose(params);
  // generate_parameters.cpp:88-108
  FILE *input = fopen(input_path, "w");
  write_fr<ppT>(input, Fr<ppT>::one());
  for (size_t i = 0; i < m; ++i) write_fr<ppT>(input, full[i]);
  for (size_t i = 0; i < d_plus_1; ++i) write_fr<ppT>(input, ca[i]);
  for (size_t i = 0; i < d_plus_1; ++i) write_fr<ppT>(input, cb[i]);
  for (size_t i = 0; i < d_plus_1; ++i) write_fr<ppT>(input, cc[i]);
  write_fr<ppT>(input, Fr<ppT>::random_element());
  fclose(input);
  // what the parameter file leaves out
  FILE *extras = fopen(extras_path, "w");
  write_g1<ppT>(extras, keypair.pk.alpha_g1);
  write_g1<ppT>(extras, keypair.pk.beta_g1);
  write_g1<ppT>(extras, keypair.pk.delta_g1);
  write_g2<ppT>(extras, keypair.pk.beta_g2);
  write_g2<ppT>(extras, keypair.pk.delta_g2);
  fclose(extras);
  std::ofstream vk(vk_path);
  vk << keypair.vk;
  vk.close();
  return 0;
}

template <typename ppT>
static int verify(const char *vk_path, const char *input_path, const char *proof_path) {
  ppT::init_public_params();
  libff::inhibit_profiling_info = true;
  libff::inhibit_profiling_counters = true;
  r1cs_gg_ppzksnark_verification_key<ppT> vk;
  std::ifstream vkf(vk_path);
  if (!vkf) {
    fprintf(stderr, "groth16_tool: cannot open %s\n", vk_path);
    return 2;
  }
  vkf >> vk;
  FILE *input = fopen(input_path, "r");
  FILE *pf = fopen(proof_path, "r");
  if (!input || !pf) {
    fprintf(stderr, "groth16_tool: cannot open input or proof\n");
    return 2;
  }
  read_fr<ppT>(input);  // w[0] = 1
  std::vector<Fr<ppT>> primary_input(1, read_fr<ppT>(input));  // main.cpp:302
  fclose(input);
  G1<ppT> A = read_g1<ppT>(pf);
  G2<ppT> B = read_g2<ppT>(pf);
  G1<ppT> C = read_g1<ppT>(pf);
  fclose(pf);
  r1cs_gg_ppzksnark_proof<ppT> proof(std::move(A), std::move(B), std::move(C));
  const bool ok = r1cs_gg_ppzksnark_verifier_strong_IC<ppT>(vk, primary_input, proof);
  printf("%s\n", ok ? "PROOF VERIFIES" : "PROOF REJECTED");
  return ok ? 0 : 1;
}

int main(int argc, char **argv) {
  setbuf(stdout, NULL);
  if (argc < 3) {
    fprintf(stderr, "usage: groth16_tool gen <curve> <log2> <params> <input> <extras> <vk> | verify <curve> <vk> <input> <proof>\n");
    return 2;
  }
  const std::string mode(argv[1]), curve(argv[2]);
  if (mode == "gen" && argc == 8) {
    const int k = atoi(argv[3]);
    if (curve == "MNT4753") return gen<mnt4753_pp>(k, argv[4], argv[5], argv[6], argv[7]);
    if (curve == "MNT6753") return gen<mnt6753_pp>(k, argv[4], argv[5], argv[6], argv[7]);
  }
  if (mode == "verify" && argc == 6) {
    if (curve == "MNT4753") return verify<mnt4753_pp>(argv[3], argv[4], argv[5]);
    if (curve == "MNT6753") return verify<mnt6753_pp>(argv[3], argv[4], argv[5]);
  }
  fprintf(stderr, "groth16_tool: bad arguments\n");
  return 2;
}
