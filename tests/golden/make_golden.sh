#!/usr/bin/env bash
# Mint the committed golden fixtures from the UNMODIFIED reference (needs /root/reference; run in the build container).
#   params/input : reference generator (libsnark/generate_parameters.cpp:23-123) at log2(d+1) = 5 and 8
#   output       : reference CPU prover `main` (libsnark/main.cpp) on those files
# generate_parameters is seeded from time()/urandom, so re-running produces different (equally valid) fixtures.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="$HERE/../../oracle/_ref"
bash "$HERE/../../oracle/build_ref.sh" >/dev/null
cd "$HERE"
for curve in MNT4753 MNT6753; do
  for k in 5 8; do
    "$REF/gen_params_any" $curve $k ${curve}_k$k.params ${curve}_k$k.input >/dev/null
    "$REF/main" $curve compute ${curve}_k$k.params ${curve}_k$k.input ${curve}_k$k.output >/dev/null
    "$REF/piecewise_host" $curve compute ${curve}_k$k.params ${curve}_k$k.input /tmp/pw_$$.out >/dev/null 2>&1
    cmp ${curve}_k$k.output /tmp/pw_$$.out   # main (Bos-Coster) == piecewise (BDLO12 through B::)
    rm -f /tmp/pw_$$.out
  done
done
sha256sum *.output > SHA256SUMS
cat SHA256SUMS
