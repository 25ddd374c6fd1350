// seaweedfs_b200/csrc/codegen.h — straight-line GF(2^8) matrix-apply generator.
//
// Given an R×K coefficient matrix over GF(2^8)/0x11D, emits a CUDA struct with a static device
// function  combine(const u32 (&x)[K], u32 (&y)[R])  that computes, on four packed byte columns at a time (one 32-bit word per stream),
//     y[p] = XOR_i  M[p][i] ⊗ x[i]
// i.e. exactly what reedsolomon.Encoder.Encode / Reconstruct compute per byte column
// (call sites weed/storage/erasure_coding/ec_encoder.go:265,360; in-tree statement
// seaweed-volume/vendor/reed-solomon-erasure/src/core.rs:484-512).
//
// Formulation: bit-plane Horner.  Write each coefficient as Σ_b c_b·2^b; then
//     y = Σ_b 2^b · S_b,   S_b = XOR of the inputs whose coefficient has bit b set,
// evaluated as y = (((S_7·2 ^ S_6)·2 ^ S_5)·2 …) with a SWAR "multiply four packed bytes by 2"
// step.  No tables, no shared memory, no data-dependent addressing: ~2 LOP3 + 3 IMAD per step
// and one 3-input LOP3 per two XOR terms.  Two optimisations reduce the instruction count:
//   * output basis: instead of the R rows, evaluate R GF(2)-independent XOR-combinations of rows
//     whose coefficients have fewer/lower bits (e.g. row10^row11 of RS(10,4) only has degree 4),
//     then recover the rows with a few XORs;
//   * common-subexpression extraction over the S_b sets (greedy pairs/triples, 3-input XOR cost).
// The same generator feeds the ahead-of-time RS(10,4) encode kernel (build time) and the
// NVRTC-specialised reconstruct kernels (run time).
#pragma once
#include <string>
#include <vector>

#include "gf256.h"

namespace swec {

struct CodegenStats {
    int xtime_steps = 0;  // SWAR multiply-by-2 steps
    int xor_ops = 0;      // 2-/3-input XOR instructions
    int shared_signals = 0;
    int terms_before = 0, terms_after = 0;
};

struct CodegenOptions {
    bool optimise_basis = true;
    bool extract_common = true;
    // input-side power chains for signals that alone occupy the high bit-planes of several rows (codegen.cc 3b).
    // Off by default: the kernels it produces have been verified on the CPU only (tests/test_codegen.py), not yet
    // measured on a B200 — "encode_formulation" / SWEC_ENCODE_FORMULATION=1 selects the RS(10,4) encode kernel built with it.
    bool share_powers = false;
    // explicit output basis: R masks over the rows (bit p = row p takes part), GF(2)-independent; empty = choose one.
    // Used by `swec_codegen --search-basis`, which tries every basis under the full cost (steps AND XORs after CSE and
    // power sharing) instead of the a-priori estimate the built-in choice uses.
    std::vector<unsigned> basis;
};

// Returns CUDA source text defining
//   struct <struct_name> { static constexpr int K, R; __device__ static void combine(x, y); };
std::string generate_combine(const Matrix& rows, const std::string& struct_name,
                             const CodegenOptions& opt, CodegenStats* stats);

}  // namespace swec
