#!/usr/bin/env bash
# Oracle infrastructure (NOT product code).
# Compiles the UNMODIFIED reference CPU prover from the sources where they lie under /root/reference into
# oracle/_ref/ (git-ignored, travels to the GPU box with gpurun). Nothing is copied into the repo.
#
#   oracle/_ref/main                 <- libsnark/main.cpp               (the sha256 oracle, Bos-Coster, OpenMP)
#   oracle/_ref/generate_parameters  <- libsnark/generate_parameters.cpp
#   oracle/_ref/gen_params_any       <- oracle/gen_params_any.cpp (includes the reference generator, any size)
#   oracle/_ref/piecewise_host       <- cuda_prover_piecewise.cu + libsnark/prover_reference_functions.cpp
#                                        (host-only build of the reference's B:: path, BDLO12)
#
# The reference's own CMake build is not used (needs gmp.h, boost, procps: absent here). Recipe = SURVEY.md 8c.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
R="${REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$R/libsnark" ]; then
  echo "build_ref.sh: $R not present (GPU box?) - using prebuilt oracle/_ref if any" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/stub/array" "$OUT/stub/fixnum" "$OUT/stub/functions" "$OUT/stub/modnum"
GMPSO="$(ls /usr/lib/x86_64-linux-gnu/libgmp.so.10 2>/dev/null || true)"
[ -n "$GMPSO" ] || { echo "libgmp.so.10 missing" >&2; exit 1; }
ln -sf "$GMPSO" "$OUT/libgmp.so"
FLAGS="-std=c++14 -O2 -fopenmp -w -DCURVE_MNT4753 -DMULTICORE=1 -DBINARY_OUTPUT -DMONTGOMERY_OUTPUT \
 -DNO_PT_COMPRESSION=1 -DNO_PROCPS -DUSE_ASM -I$HERE/shim -I$R -I$R/depends/libff -I$R/depends/libfqfft"
LF="$R/depends/libff/libff"
SRCS="algebra/curves/mnt753/mnt46753_common.cpp
 algebra/curves/mnt753/mnt4753/mnt4753_g1.cpp algebra/curves/mnt753/mnt4753/mnt4753_g2.cpp
 algebra/curves/mnt753/mnt4753/mnt4753_init.cpp algebra/curves/mnt753/mnt4753/mnt4753_pairing.cpp
 algebra/curves/mnt753/mnt4753/mnt4753_pp.cpp
 algebra/curves/mnt753/mnt6753/mnt6753_g1.cpp algebra/curves/mnt753/mnt6753/mnt6753_g2.cpp
 algebra/curves/mnt753/mnt6753/mnt6753_init.cpp algebra/curves/mnt753/mnt6753/mnt6753_pairing.cpp
 algebra/curves/mnt753/mnt6753/mnt6753_pp.cpp
 common/double.cpp common/profiling.cpp common/utils.cpp"
pids=()
for s in $SRCS; do
  o="$OUT/obj/$(basename "${s%.cpp}").o"
  if [ ! -f "$o" ]; then g++ $FLAGS -c "$LF/$s" -o "$o" & pids+=($!); fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
LIBS="-L$OUT -lgmp -lcrypto"
build() { # name, sources...
  local name="$1"; shift
  if [ ! -x "$OUT/$name" ]; then g++ $FLAGS "$@" "$OUT"/obj/*.o $LIBS -o "$OUT/$name"; fi
}
build main "$R/libsnark/main.cpp" &
build generate_parameters "$R/libsnark/generate_parameters.cpp" &
build gen_params_any "$HERE/gen_params_any.cpp" &
build groth16_tool "$HERE/groth16_tool.cpp" &
# host-only build of the reference's own piecewise driver: its six cuda-fixnum includes are satisfied by empty
# stub headers (it calls none of them), and -x c++ treats the .cu as plain C++.
for h in array/fixnum_array.h fixnum/warp_fixnum.cu functions/modexp.cu functions/multi_modexp.cu \
         modnum/modnum_monty_cios.cu modnum/modnum_monty_redc.cu; do : > "$OUT/stub/$h"; done
if [ ! -x "$OUT/piecewise_host" ]; then
  g++ $FLAGS -I"$R/libsnark/prover_reference_include" -I"$OUT/stub" \
     -x c++ "$R/cuda_prover_piecewise.cu" -x c++ "$R/libsnark/prover_reference_functions.cpp" -x none \
     "$OUT"/obj/*.o $LIBS -o "$OUT/piecewise_host" &
fi
wait
# Drop-in check of the boundary: the reference's OWN, unmodified prover driver (cuda_prover_piecewise.cu) compiled
# against THIS repo's prover_reference_functions.hpp and linked to the B200 library instead of libff.
PKG="$HERE/../snark_challenge_prover_reference_b200"
if [ -f "$PKG/libb200groth16.so" ]; then
  if [ ! -x "$OUT/piecewise_b200" ] || [ "$PKG/libb200groth16.so" -nt "$OUT/piecewise_b200" ]; then
    g++ -std=c++14 -O2 -w -I"$PKG/csrc/host" -I"$HERE/../include" -I"$OUT/stub" \
       -x c++ "$R/cuda_prover_piecewise.cu" -x c++ "$PKG/csrc/host/b200_bundle.cpp" -x none \
       -L"$PKG" -lb200groth16 -Wl,-rpath,"$PKG" -Wl,-rpath,'$ORIGIN/../../snark_challenge_prover_reference_b200' \
       -o "$OUT/piecewise_b200"
  fi
fi
ls -la "$OUT" | grep -v obj
