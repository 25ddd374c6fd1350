// seaweedfs_b200/csrc/gf256.cc — see gf256.h.
#include "gf256.h"

#include <algorithm>

namespace swec {

GF::GF() {
    // carry-less shift-and-add multiplication reduced by 0x11D
    for (unsigned a = 0; a < 256; a++)
        for (unsigned b = 0; b < 256; b++) {
            unsigned acc = 0, x = a;
            for (unsigned bit = 0; bit < 8; bit++) {
                if (b & (1u << bit)) acc ^= x;
                x <<= 1;
                if (x & 0x100) x ^= kFieldPoly;
            }
            mul[a][b] = uint8_t(acc);
        }
    inv[0] = 0;
    for (unsigned a = 1; a < 256; a++)
        for (unsigned b = 1; b < 256; b++)
            if (mul[a][b] == 1) { inv[a] = uint8_t(b); break; }
}

const GF& GF::get() {
    static const GF g;
    return g;
}

uint8_t GF::pow(uint8_t a, unsigned n) const {
    uint8_t r = 1;  // a^0 = 1 also for a = 0, as the reference's exp() (galois_8.rs:89-103)
    for (unsigned i = 0; i < n; i++) r = mul[r][a];
    return r;
}

Matrix mat_identity(int n) {
    Matrix m(n, n);
    for (int i = 0; i < n; i++) m.at(i, i) = 1;
    return m;
}

Matrix mat_mul(const Matrix& a, const Matrix& b) {
    const GF& gf = GF::get();
    Matrix out(a.rows, b.cols);
    for (int r = 0; r < a.rows; r++)
        for (int c = 0; c < b.cols; c++) {
            uint8_t acc = 0;
            for (int i = 0; i < a.cols; i++) acc ^= gf.mul[a.at(r, i)][b.at(i, c)];
            out.at(r, c) = acc;
        }
    return out;
}

bool mat_inv(const Matrix& m, Matrix* out) {
    const GF& gf = GF::get();
    const int n = m.rows;
    if (m.cols != n) return false;
    Matrix a = m, b = mat_identity(n);
    for (int col = 0; col < n; col++) {
        int piv = col;
        while (piv < n && a.at(piv, col) == 0) piv++;
        if (piv == n) return false;
        if (piv != col)
            for (int c = 0; c < n; c++) {
                std::swap(a.at(piv, c), a.at(col, c));
                std::swap(b.at(piv, c), b.at(col, c));
            }
        const uint8_t s = gf.inv[a.at(col, col)];
        for (int c = 0; c < n; c++) {
            a.at(col, c) = gf.mul[s][a.at(col, c)];
            b.at(col, c) = gf.mul[s][b.at(col, c)];
        }
        for (int r = 0; r < n; r++) {
            if (r == col) continue;
            const uint8_t f = a.at(r, col);
            if (!f) continue;
            for (int c = 0; c < n; c++) {
                a.at(r, c) ^= gf.mul[f][a.at(col, c)];
                b.at(r, c) ^= gf.mul[f][b.at(col, c)];
            }
        }
    }
    *out = b;
    return true;
}

Matrix rs_generator(int k, int m) {
    const GF& gf = GF::get();
    const int total = k + m;
    Matrix vand(total, k), top(k, k), top_inv;
    for (int r = 0; r < total; r++)
        for (int c = 0; c < k; c++) vand.at(r, c) = gf.pow(uint8_t(r), unsigned(c));
    for (int r = 0; r < k; r++)
        for (int c = 0; c < k; c++) top.at(r, c) = vand.at(r, c);
    mat_inv(top, &top_inv);  // distinct evaluation points ⇒ always invertible
    return mat_mul(vand, top_inv);
}

bool rs_reconstruct_plan(const Matrix& gen, int k, const uint8_t* present, bool data_only,
                         std::vector<int>* inputs, std::vector<int>* outputs, Matrix* fused) {
    const int total = gen.rows;
    inputs->clear();
    outputs->clear();
    for (int i = 0; i < total && int(inputs->size()) < k; i++)
        if (present[i]) inputs->push_back(i);
    if (int(inputs->size()) < k) return false;
    for (int i = 0; i < total; i++)
        if (!present[i] && (i < k || !data_only)) outputs->push_back(i);
    Matrix sub(k, k), dec;
    for (int r = 0; r < k; r++)
        for (int c = 0; c < k; c++) sub.at(r, c) = gen.at((*inputs)[r], c);
    if (!mat_inv(sub, &dec)) return false;
    *fused = Matrix(int(outputs->size()), k);
    for (size_t o = 0; o < outputs->size(); o++) {
        const int idx = (*outputs)[o];
        if (idx < k) {
            for (int c = 0; c < k; c++) fused->at(int(o), c) = dec.at(idx, c);
        } else {
            Matrix prow(1, k);
            for (int c = 0; c < k; c++) prow.at(0, c) = gen.at(idx, c);
            Matrix f = mat_mul(prow, dec);
            for (int c = 0; c < k; c++) fused->at(int(o), c) = f.at(0, c);
        }
    }
    return true;
}

}  // namespace swec
