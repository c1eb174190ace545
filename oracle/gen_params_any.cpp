// Oracle infrastructure (NOT product code): drives the UNMODIFIED reference parameter generator at an
// arbitrary domain size. The reference's own main() (libsnark/generate_parameters.cpp:125-137) hard-codes
// log2(d+1) in {20,15} / `fast` {14,10}; its template generate_paramaters<ppT>(log2_d, params, input)
// (generate_parameters.cpp:23-123) accepts any size. We include the reference source where it lies and only
// rename its main so that small, committed fixtures (tests/golden/) can be minted.
//
//   usage: gen_params_any <MNT4753|MNT6753> <log2_d_plus_1> <params_out> <input_out>
#define main reference_generate_parameters_main
#include <libsnark/generate_parameters.cpp>
#undef main

int main(int argc, char **argv) {
  if (argc != 5) {
    fprintf(stderr, "usage: %s <MNT4753|MNT6753> <log2_d_plus_1> <params_out> <input_out>\n", argv[0]);
    return 2;
  }
  std::string curve(argv[1]);
  int k = atoi(argv[2]);
  if (curve == "MNT4753") return generate_paramaters<mnt4753_pp>(k, argv[3], argv[4]);
  if (curve == "MNT6753") return generate_paramaters<mnt6753_pp>(k, argv[3], argv[4]);
  fprintf(stderr, "unknown curve %s\n", argv[1]);
  return 2;
}
