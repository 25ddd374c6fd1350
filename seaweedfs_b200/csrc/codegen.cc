// seaweedfs_b200/csrc/codegen.cc — see codegen.h.
#include "codegen.h"

#include <algorithm>
#include <array>
#include <cstdio>
#include <map>
#include <sstream>
#include <vector>

namespace swec {
namespace {

int top_bit(unsigned v) {
    int b = -1;
    while (v) { b++; v >>= 1; }
    return b;
}

struct VRow {
    std::vector<uint8_t> coef;  // K coefficients
    int degree() const {
        unsigned all = 0;
        for (uint8_t c : coef) all |= c;
        return top_bit(all);
    }
    int terms() const {
        int t = 0;
        for (uint8_t c : coef) t += __builtin_popcount(c);
        return t;
    }
    // relative issue cost: a multiply-by-2 step ≈ 5 instructions, a term ≈ half a LOP3
    double cost() const { return 5.0 * std::max(degree(), 0) + 0.5 * terms(); }
};

VRow combine_rows(const Matrix& m, unsigned mask) {
    VRow v;
    v.coef.assign(size_t(m.cols), 0);
    for (int p = 0; p < m.rows; p++)
        if (mask & (1u << p))
            for (int i = 0; i < m.cols; i++) v.coef[size_t(i)] ^= m.at(p, i);
    return v;
}

// rank over GF(2) of R masks of R bits
bool independent(const std::vector<unsigned>& masks) {
    std::vector<unsigned> basis;
    for (unsigned m : masks) {
        for (unsigned b : basis) m = std::min(m, m ^ b);
        if (!m) return false;
        basis.push_back(m);
    }
    return true;
}

// Solve y_p = XOR_{q in sel[p]} v_q given v_q = XOR_{p in masks[q]} y_p  (invert over GF(2)).
std::vector<unsigned> invert_gf2(const std::vector<unsigned>& masks) {
    const int n = int(masks.size());
    std::vector<unsigned> a = masks, b(size_t(n), 0u);
    for (int i = 0; i < n; i++) b[size_t(i)] = 1u << i;
    // rows of a: v_q expressed in y; row-reduce [a | b] to [I | a^-1]: then y_p = XOR of v_q in b'
    for (int col = 0; col < n; col++) {
        int piv = col;
        while (piv < n && !(a[size_t(piv)] & (1u << col))) piv++;
        std::swap(a[size_t(piv)], a[size_t(col)]);
        std::swap(b[size_t(piv)], b[size_t(col)]);
        for (int r = 0; r < n; r++)
            if (r != col && (a[size_t(r)] & (1u << col))) {
                a[size_t(r)] ^= a[size_t(col)];
                b[size_t(r)] ^= b[size_t(col)];
            }
    }
    return b;  // b[p] = mask over q of the virtual rows that XOR to y_p
}

struct Emitter {
    std::ostringstream os;
    int next_tmp = 0;
    CodegenStats* st;
    std::string tmp() { return "t" + std::to_string(next_tmp++); }
    // XOR a list of named values into at most `keep` values, using 3-input XORs
    std::string xor_all(std::vector<std::string> v) {
        if (v.empty()) return "0u";
        while (v.size() > 1) {
            std::string t = tmp();
            if (v.size() >= 3) {
                os << "    const u32 " << t << " = SWEC_X3(" << v[0] << ", " << v[1] << ", " << v[2] << ");\n";
                v.erase(v.begin(), v.begin() + 3);
            } else {
                os << "    const u32 " << t << " = SWEC_X2(" << v[0] << ", " << v[1] << ");\n";
                v.erase(v.begin(), v.begin() + 2);
            }
            st->xor_ops++;
            v.push_back(t);
        }
        return v[0];
    }
};

}  // namespace

std::string generate_combine(const Matrix& rows, const std::string& struct_name,
                             const CodegenOptions& opt, CodegenStats* stats_out) {
    CodegenStats stats;
    const int R = rows.rows, K = rows.cols;

    // ---- 1. output basis ------------------------------------------------------------------
    std::vector<unsigned> masks(static_cast<size_t>(R));
    for (int p = 0; p < R; p++) masks[size_t(p)] = 1u << p;
    if (int(opt.basis.size()) == R && independent(opt.basis)) {
        masks = opt.basis;
    } else if (opt.optimise_basis && R >= 2 && R <= 4) {
        // exhaustive: choose R independent non-zero combinations with the least total cost
        const unsigned ncomb = (1u << R) - 1;
        std::vector<double> cost(ncomb + 1, 0.0);
        for (unsigned c = 1; c <= ncomb; c++) cost[c] = combine_rows(rows, c).cost();
        double best = 1e30;
        std::vector<unsigned> pick(static_cast<size_t>(R)), cur(static_cast<size_t>(R));
        // recursive enumeration of increasing R-subsets
        std::vector<int> idx(static_cast<size_t>(R));
        for (int i = 0; i < R; i++) idx[size_t(i)] = i + 1;
        while (true) {
            for (int i = 0; i < R; i++) cur[size_t(i)] = unsigned(idx[size_t(i)]);
            if (independent(cur)) {
                double c = 0;
                for (unsigned m : cur) c += cost[m];
                // recombination XORs
                for (unsigned sel : invert_gf2(cur)) c += 0.5 * (__builtin_popcount(sel) - 1);
                if (c < best - 1e-9) { best = c; pick = cur; }
            }
            int i = R - 1;
            while (i >= 0 && idx[size_t(i)] == int(ncomb) - (R - 1 - i)) i--;
            if (i < 0) break;
            idx[size_t(i)]++;
            for (int j = i + 1; j < R; j++) idx[size_t(j)] = idx[size_t(j - 1)] + 1;
        }
        masks = pick;
    } else if (opt.optimise_basis && R > 4 && R <= 32) {
        // greedy elementary row operations v_p ^= v_q while the total cost drops
        bool improved = true;
        while (improved) {
            improved = false;
            for (int p = 0; p < R; p++)
                for (int q = 0; q < R; q++) {
                    if (p == q) continue;
                    const unsigned cand = masks[size_t(p)] ^ masks[size_t(q)];
                    if (combine_rows(rows, cand).cost() + 0.5 < combine_rows(rows, masks[size_t(p)]).cost()) {
                        masks[size_t(p)] = cand;
                        improved = true;
                    }
                }
        }
    }
    std::vector<VRow> vrows;
    for (unsigned m : masks) vrows.push_back(combine_rows(rows, m));
    const std::vector<unsigned> recombine = invert_gf2(masks);

    // ---- 2. term sets S[q][b] over signals (signal ids: 0..K-1 inputs, then shared temps) ----
    struct Set { int q, b; std::vector<int> sig; };
    std::vector<Set> sets;
    for (int q = 0; q < R; q++)
        for (int b = 0; b <= vrows[size_t(q)].degree(); b++) {
            Set s{q, b, {}};
            for (int i = 0; i < K; i++)
                if (vrows[size_t(q)].coef[size_t(i)] & (1u << b)) s.sig.push_back(i);
            stats.terms_before += int(s.sig.size());
            sets.push_back(s);
        }

    // ---- 3. greedy common-subexpression extraction (3-input XOR cost model) ----------------
    std::vector<std::array<int, 3>> shared;  // shared signal s = K + index → up to 3 operands (-1 = none)
    if (opt.extract_common) {
        while (true) {
            std::map<std::array<int, 3>, int> freq;
            for (const Set& s : sets) {
                const auto& g = s.sig;
                const int n = int(g.size());
                for (int a = 0; a < n; a++)
                    for (int b = a + 1; b < n; b++) {
                        freq[{g[size_t(a)], g[size_t(b)], -1}]++;
                        for (int c = b + 1; c < n; c++) freq[{g[size_t(a)], g[size_t(b)], g[size_t(c)]}]++;
                    }
            }
            double best_gain = 0.0;
            std::array<int, 3> best{-1, -1, -1};
            for (const auto& kv : freq) {
                const bool triple = kv.first[2] >= 0;
                // each use of a triple turns 3 terms into 1 (saves one LOP3); a pair saves half of one
                const double gain = (triple ? 1.0 : 0.5) * kv.second - 1.0;
                if (gain > best_gain + 1e-9) { best_gain = gain; best = kv.first; }
            }
            if (best[0] < 0) break;
            const int id = K + int(shared.size());
            shared.push_back(best);
            for (Set& s : sets) {
                auto has = [&](int v) { return v < 0 || std::find(s.sig.begin(), s.sig.end(), v) != s.sig.end(); };
                if (has(best[0]) && has(best[1]) && has(best[2])) {
                    for (int v : best)
                        if (v >= 0) s.sig.erase(std::find(s.sig.begin(), s.sig.end(), v));
                    s.sig.push_back(id);
                }
            }
        }
    }
    stats.shared_signals = int(shared.size());
    for (const Set& s : sets) stats.terms_after += int(s.sig.size());

    // ---- 3b. shared power chains (opt.share_powers) -----------------------------------------------
    // Horner on the OUTPUT side costs one multiply-by-2 step per degree of every row.  When the high bit-planes of
    // several rows hold nothing but one and the same signal sigma (RS(10,4): the two pair-sum rows of the basis are
    // c * (x0^...^x7) plus degree-0 terms), it is cheaper to build the powers 2^b * sigma ONCE on the input side and XOR
    // the ones each row needs: rows q1 (degree 4) and q3 (degree 5) of the encode matrix share one 5-step chain
    // instead of paying 4 + 5 steps.  Greedy over signals; a signal moves out of a row only above the degree of the
    // row's other terms (below it the term rides in a Horner step for free).
    std::vector<std::vector<std::pair<int, int>>> power_terms(static_cast<size_t>(R));  // per row: (signal, exponent)
    std::map<int, int> power_need;                                                      // signal -> longest chain
    if (opt.share_powers) {
        auto set_ref = [&](int q, int b) -> Set* {
            for (Set& s : sets)
                if (s.q == q && s.b == b) return &s;
            return nullptr;
        };
        auto row_degree = [&](int q, int without) {  // highest non-empty plane, ignoring `without` at planes >= 1
            int d = -1;
            for (const Set& s : sets) {
                if (s.q != q) continue;
                size_t n = s.sig.size();
                if (without >= 0 && s.b >= 1 && std::find(s.sig.begin(), s.sig.end(), without) != s.sig.end()) n--;
                if (n > 0) d = std::max(d, s.b);
            }
            return d;
        };
        const int nsig = K + int(shared.size());
        for (;;) {
            double best_gain = 0;
            int best_sig = -1;
            for (int sg = 0; sg < nsig; sg++) {
                int steps_saved = 0, chain = power_need.count(sg) ? power_need[sg] : 0, chain_new = chain, moved = 0;
                for (int q = 0; q < R; q++) {
                    const int before = row_degree(q, -1), after = std::max(0, row_degree(q, sg));
                    if (before <= after) continue;
                    steps_saved += before - after;
                    for (int b = after + 1; b <= before; b++) {
                        const Set* st = set_ref(q, b);
                        if (st && std::find(st->sig.begin(), st->sig.end(), sg) != st->sig.end()) {
                            moved++;
                            chain_new = std::max(chain_new, b);
                        }
                    }
                }
                // a step is ~5 instructions, an XORed power term ~half a LOP3; moved terms used to ride for free
                const double gain = 5.0 * (steps_saved - (chain_new - chain)) - 0.5 * moved;
                if (gain > best_gain + 1e-9) { best_gain = gain; best_sig = sg; }
            }
            if (best_sig < 0) break;
            for (int q = 0; q < R; q++) {
                const int before = row_degree(q, -1), after = std::max(0, row_degree(q, best_sig));
                if (before <= after) continue;
                for (int b = after + 1; b <= before; b++) {
                    Set* st = set_ref(q, b);
                    if (!st) continue;
                    auto it = std::find(st->sig.begin(), st->sig.end(), best_sig);
                    if (it == st->sig.end()) continue;
                    st->sig.erase(it);
                    power_terms[size_t(q)].push_back({best_sig, b});
                    power_need[best_sig] = std::max(power_need[best_sig], b);
                }
            }
        }
    }

    // ---- 4. emit ---------------------------------------------------------------------------
    Emitter em;
    em.st = &stats;
    auto name = [&](int sig) { return sig < K ? "x[" + std::to_string(sig) + "]" : "s" + std::to_string(sig - K); };
    em.os << "// generated by seaweedfs_b200/csrc/codegen.cc — do not edit\n";
    em.os << "// matrix (" << R << "x" << K << "):";
    for (int p = 0; p < R; p++) {
        em.os << (p ? " |" : "");
        for (int i = 0; i < K; i++) em.os << " " << int(rows.at(p, i));
    }
    em.os << "\nstruct " << struct_name << " {\n  static constexpr int K = " << K << ", R = " << R << ";\n";
    em.os << "  __device__ static __forceinline__ void combine(const u32 (&x)[" << K << "], u32 (&y)[" << R << "]) {\n";
    for (size_t i = 0; i < shared.size(); i++) {
        const auto& sh = shared[i];
        em.os << "    const u32 s" << i << " = ";
        if (sh[2] >= 0) em.os << "SWEC_X3(" << name(sh[0]) << ", " << name(sh[1]) << ", " << name(sh[2]) << ");\n";
        else em.os << "SWEC_X2(" << name(sh[0]) << ", " << name(sh[1]) << ");\n";
        stats.xor_ops++;
    }
    auto power_name = [&](int sig, int b) { return "p" + std::to_string(sig) + "_" + std::to_string(b); };
    for (const auto& kv : power_need) {  // 2^b * sigma for b = 1 .. longest use
        std::string prev = name(kv.first);
        for (int b = 1; b <= kv.second; b++) {
            stats.xtime_steps++;
            em.os << "    const u32 " << power_name(kv.first, b) << " = SWEC_XT0" << ((stats.xtime_steps & 1) ? "B" : "A") << "(" << prev << ");\n";
            prev = power_name(kv.first, b);
        }
    }
    std::vector<std::string> vname(static_cast<size_t>(R));
    for (int q = 0; q < R; q++) {
        static const Set kEmpty{0, 0, {}};
        auto set_of = [&](int b) -> const Set& {
            for (const Set& s : sets)
                if (s.q == q && s.b == b) return s;
            return kEmpty;
        };
        int deg = -1;  // highest plane that still has terms (shared power chains may have emptied the top ones)
        for (const Set& s : sets)
            if (s.q == q && !s.sig.empty()) deg = std::max(deg, s.b);
        std::vector<std::string> extra;  // powers of shared signals this row takes from the input-side chains
        for (const auto& pt : power_terms[size_t(q)]) extra.push_back(power_name(pt.first, pt.second));
        if (deg < 0) {
            vname[size_t(q)] = extra.empty() ? "0u" : em.xor_all(extra);
            continue;
        }
        std::vector<std::string> terms;
        for (int sg : set_of(deg).sig) terms.push_back(name(sg));
        std::string acc = em.xor_all(terms);
        for (int b = deg - 1; b >= 0; b--) {
            terms.clear();
            for (int sg : set_of(b).sig) terms.push_back(name(sg));
            std::string nxt = em.tmp();
            stats.xtime_steps++;
            if (terms.empty()) {
                em.os << "    const u32 " << nxt << " = SWEC_XT0" << ((stats.xtime_steps & 1) ? "B" : "A") << "(" << acc << ");\n";
            } else {
                // first term rides in the step's final 3-input XOR; the rest are folded off the
                // accumulator's dependency chain first
                std::string first = terms[0];
                terms.erase(terms.begin());
                std::string side;
                if (terms.size() > 2) {
                    std::vector<std::string> head(terms.begin(), terms.end() - 1);
                    std::string folded = em.xor_all(head);
                    terms = {folded, terms.back()};
                }
                // steps alternate between the A and B spelling so that a build may give them different
                // instruction mixes (device_common.cuh, SWEC_XT_VARIANT 3); otherwise A == B
                em.os << "    const u32 " << nxt << "_ = SWEC_XT1" << ((stats.xtime_steps & 1) ? "B" : "A") << "(" << acc << ", " << first << ");\n";
                if (terms.empty()) {
                    em.os << "    const u32 " << nxt << " = " << nxt << "_;\n";
                } else if (terms.size() == 1) {
                    em.os << "    const u32 " << nxt << " = SWEC_X2(" << nxt << "_, " << terms[0] << ");\n";
                    stats.xor_ops++;
                } else {
                    em.os << "    const u32 " << nxt << " = SWEC_X3(" << nxt << "_, " << terms[0] << ", " << terms[1] << ");\n";
                    stats.xor_ops++;
                }
            }
            acc = nxt;
        }
        if (!extra.empty()) {
            extra.insert(extra.begin(), acc);
            acc = em.xor_all(extra);
        }
        vname[size_t(q)] = acc;
    }
    for (int p = 0; p < R; p++) {
        std::vector<std::string> parts;
        for (int q = 0; q < R; q++)
            if (recombine[size_t(p)] & (1u << q)) parts.push_back(vname[size_t(q)]);
        std::string v = em.xor_all(parts);
        em.os << "    y[" << p << "] = " << v << ";\n";
    }
    em.os << "  }\n};\n";
    em.os << "// stats: xtime_steps=" << stats.xtime_steps << " xor_ops=" << stats.xor_ops
          << " shared=" << stats.shared_signals << " terms " << stats.terms_before << "->" << stats.terms_after << "\n";
    if (stats_out) *stats_out = stats;
    return em.os.str();
}

}  // namespace swec
