// Drop-in name: the reference's driver does `#include <prover_reference_functions.hpp>` (cuda_prover_piecewise.cu:3).
// The B200-backed bundle lives in b200_bundle.hpp.
#pragma once
#include "b200_bundle.hpp"
