// seaweedfs_b200/csrc/gf256.h — host-side GF(2^8) arithmetic and coding matrices.
//
// Field and matrix construction must match what SeaweedFS gets from
// reedsolomon.New(dataShards, parityShards) with no options
// (weed/storage/erasure_coding/ec_context.go:34-36): GF(2^8) over x^8+x^4+x^3+x^2+1 (0x11D),
// generator matrix = Vandermonde(total×k, entry r^c) · inverse(top k×k)
// (in-tree statement: seaweed-volume/vendor/reed-solomon-erasure/src/core.rs:431-437,
//  src/matrix.rs:263-276).  Only the resulting bytes matter; the algorithms here are our own.
#pragma once
#include <cstdint>
#include <vector>

namespace swec {

constexpr unsigned kFieldPoly = 0x11D;

struct GF {
    uint8_t mul[256][256];
    uint8_t inv[256];
    GF();
    static const GF& get();
    uint8_t pow(uint8_t a, unsigned n) const;
};

// Row-major byte matrix.
struct Matrix {
    int rows = 0, cols = 0;
    std::vector<uint8_t> v;
    Matrix() = default;
    Matrix(int r, int c) : rows(r), cols(c), v(size_t(r) * c, 0) {}
    uint8_t& at(int r, int c) { return v[size_t(r) * cols + c]; }
    uint8_t at(int r, int c) const { return v[size_t(r) * cols + c]; }
    const uint8_t* row(int r) const { return &v[size_t(r) * cols]; }
    bool operator==(const Matrix& o) const { return rows == o.rows && cols == o.cols && v == o.v; }
};

Matrix mat_mul(const Matrix& a, const Matrix& b);
bool mat_inv(const Matrix& m, Matrix* out);  // false if singular
Matrix mat_identity(int n);

// (k+m)×k generator: identity on top, parity rows below.
Matrix rs_generator(int k, int m);

// For a presence mask over k+m shards: the first k present shards (index order) are the inputs
// (Encoder.Reconstruct semantics, weed/storage/erasure_coding/ec_encoder.go:360; in-tree
// statement core.rs:736-926).  Produces ONE fused matrix (rows = every missing shard wanted,
// cols = the k inputs): decode rows for missing data, parity_row·decode for missing parity.
// Returns false if fewer than k shards are present.
bool rs_reconstruct_plan(const Matrix& gen, int k, const uint8_t* present, bool data_only,
                         std::vector<int>* inputs, std::vector<int>* outputs, Matrix* fused);

}  // namespace swec
