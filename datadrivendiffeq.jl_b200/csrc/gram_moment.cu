// Moment schedule of the fused library + Gram stage: the Gram matrix of a MONOMIAL library is a moment matrix.
//
// G[t, u] = sum_s theta_t(s) theta_u(s) only depends on the product monomial theta_t * theta_u, i.e. on the
// multiset union of the two terms' factors.  A polynomial library of total degree D over n states (reference
// src/utils/basis_generators.jl:84-126) has k(k+1)/2 Gram entries but only C(n + 2D, 2D) distinct moments: at
// the Lorenz-96 configuration (n = 64, D = 2, k = 2145) that is 814 385 moments for 2 301 585 entries.  Every
// moment is computed ONCE here, on the same FP64 DMMA kernel (gram.cu) as the plain schedule:
//
//   * a moment's 2D factors (padded with the constant 1 as variable 0), sorted, are cut in the middle:
//     P = the D smallest, Q = the D largest, so that max(P) <= min(Q);
//   * the distinct P's, sorted by max(P), are the ROWS of a virtual library, the distinct Q's, sorted by
//     decreasing min(Q), its COLUMNS (the targets come first among the columns: R = Theta Y' needs every
//     row; they are also appended to the rows for diag(Y Y'));
//   * the needed (P, Q) entries then form a staircase in the rows x columns plane, and only the 128 x 128
//     block pairs that the staircase touches are contracted: 78 at Lorenz-96 against 153 + 18 diagonal, and
//     pairs at its edge that only need their first 64 rows (or columns: run transposed) use the kernel's
//     64 x 128 patch;
//   * a gather kernel writes every entry of the packed [G | R | yy] from the tile entry of its moment
//     (G comes out exactly symmetric, and bit-reproducible: fixed slot order, no atomics).
//
// Nothing about the data path changes: X / Y tiles arrive by TMA, the (virtual) terms are multiplied out in
// shared memory, DMMA.8x8x4 accumulates in registers; Theta is never materialised.  The CTAs run the block
// pairs in lock-step waves (pairs x sample ranges) so that concurrently running CTAs read the same X / Y
// tiles and HBM sees every tile about once per wave.
//
// Libraries that are not closed under the split (a P or Q that is not itself cheap to add) still work: the
// virtual rows / columns are whatever monomials the splits produce.  The schedule is used when it needs
// fewer block pairs than the plain one (see `eligible` below); DDB_GRAM_MOMENT=0 disables it.
#include "common.cuh"
#include "mma_common.cuh"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace ddb {

int gram_launch_bt128(ddb_ctx* ctx, const GramArgs& args, int nctas, bool offdiag_only);
int gram_launch_mb128(ddb_ctx* ctx, const GramArgs& args, int nctas);
bool gram_mb_fits(int n, int n_t);
int gram_ensure_factors(ddb_ctx* ctx, int kaug_pad);

namespace {

constexpr int kBT = 128;
constexpr int kMaxMomentTerms = 8192;  // the analysis walks k (k + 1) / 2 entries on the host (1.4 s at k = 5151)
constexpr double kHalfCost = 0.53;     // measured cost of a 64 x 128 patch relative to a whole 128 x 128 pair
// every wave is one more pass over the X / Y stream and one more pipeline fill per CTA: a wave must save this
// fraction of the perfectly split time to be worth it
constexpr double kWavePenaltyFrac = 0.006;
constexpr double kQuarterCost = 0.45;  // 32 x 128 patch (eight DMMAs per warp and k-step do not fill the pipe)
using Key = std::array<uint16_t, kMaxFactors>;  // factor codes, ascending, leading zeros = the constant 1
struct KeyHash {
    size_t operator()(const Key& k) const {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < kMaxFactors; ++i) h = (h ^ k[i]) * 0x100000001B3ull + (h >> 29);
        return (size_t)h;
    }
};

// sum over the slots of the entry's block pair, in slot order
__global__ void __launch_bounds__(256) moment_gather_kernel(const double* __restrict__ partials,
                                                            const int32_t* __restrict__ pair_slot_begin,
                                                            const int32_t* __restrict__ src, int64_t count,
                                                            double* __restrict__ packed) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    const int32_t sv = src[e];
    const int pair = sv >> 14, off = sv & (kBT * kBT - 1);
    double s = 0.0;
    for (int sl = pair_slot_begin[pair]; sl < pair_slot_begin[pair + 1]; ++sl)
        s += partials[(size_t)sl * (kBT * kBT) + off];
    packed[e] = s;
}

// cost of a pair that needs its first `rows` rows, relative to a whole pair; class bits of GramSegment::pad
inline double pair_cost(int rows) { return rows == 128 ? 1.0 : rows == 64 ? kHalfCost : kQuarterCost; }
inline int pair_class(int rows) { return rows == 128 ? 0 : rows == 64 ? 1 : 2; }

}  // namespace

struct MomentState {
    // library-dependent part
    uint64_t lib_version = ~0ull;
    int k = -1, n_t = -1;
    bool eligible = false;
    int nrb = 0, ncb = 0, nvirt = 0, npairs = 0;
    std::vector<std::pair<int, int>> pairs;  // (row block, nrb + column block), whole pairs first
    std::vector<uint8_t> rows_used;          // per pair: 128, or 64 / 32 leading rows of the first block (sorted descending)
    DevBuf factors, src;
    bool mb = false;      // tensor-pipe builder tables present (degree <= 2, every pair within the kernel's slots)
    DevBuf optab, opcnt;  // [npairs][kMbPairWords] uint32, [npairs][8] uint8
    // sample-count-dependent part: a few wave plans are kept (the chunked upload of ddb_upload_gram walks through
    // four chunk lengths on every call: two short first chunks, the regular one, the last one)
    struct DevPlan {
        int64_t m = -1;
        int nctas = -1;
        int nslots = 0;
        DevBuf buf;
        size_t off_cta = 0, off_psb = 0;
        uint64_t stamp = 0;
    };
    static constexpr int kPlans = 6;
    DevPlan plans[kPlans];
    uint64_t clock = 0;
    DevBuf partials;
};

// Host-only analysis of a monomial library (n x k exponent table, n_t targets): virtual rows / columns, needed
// block pairs, gather map and virtual factor table.  `worth` = the schedule needs clearly fewer block pairs
// than the plain one.
struct MomentMap {
    bool worth = false;
    int nrows = 0, ncols = 0, nrb = 0, ncb = 0, nvirt = 0;
    std::vector<std::pair<int, int>> pairs;  // (first, second operand block of the kernel): (row block, nrb + column
                                             // block), or the other way round for a column-half pair; whole pairs first
    std::vector<uint8_t> rows_used;          // 128, or 64 / 32: only that many leading rows of the first operand block are needed
    std::vector<int32_t> src;                // per packed entry: (pair << 14) | (row-in-tile * 128 + column-in-tile)
    std::vector<uint16_t> tab;               // (nrb + ncb) * 128 virtual terms x kMaxFactors factor codes
    // tensor-pipe builder (gram_mb_kernel): empty when the library does not qualify
    std::vector<uint32_t> optab;             // npairs x kMbPairWords
    std::vector<uint8_t> opcnt;              // npairs x 8: bits 0..2 = k-step kk runs a third op, bits 4..7 = copy slots
    std::vector<int> block_ops;              // ops per virtual block (diagnostics)
};

// ---------------------------------------------------------------------------------------------
// Tensor-pipe builder: cover of a block's terms by (2 second factors) x (8 first factors) rectangles
// ---------------------------------------------------------------------------------------------
// Factor ids: 0 = the constant (the slot's ones column), 1..n = states, n+1..n+n_t = targets, -1 = nothing (0.0).
struct MbOp {
    int u[2];
    int v[8];
    int dst[2][8];  // tile column of u[r] * v[j], -1 = not needed
};

// Greedy cover of the term graph (vertices = factors, edges = terms {f, g}, loops = squares) by K(2,8) bicliques:
// the vertex with the most uncovered terms, the partner and the eight neighbours that cover most.
static void cover_block(const std::vector<std::array<int, 3>>& terms /* column, f, g */, std::vector<MbOp>* ops) {
    ops->clear();
    std::vector<std::array<int, 3>> left = terms;
    auto other = [](const std::array<int, 3>& e, int u) { return e[1] == u ? e[2] : e[1]; };
    while (!left.empty()) {
        std::vector<int> verts;
        for (const auto& e : left) {
            verts.push_back(e[1]);
            verts.push_back(e[2]);
        }
        std::sort(verts.begin(), verts.end());
        verts.erase(std::unique(verts.begin(), verts.end()), verts.end());
        auto neighbours = [&](int u) {
            std::vector<int> nb;
            for (const auto& e : left)
                if (e[1] == u || e[2] == u) nb.push_back(other(e, u));
            std::sort(nb.begin(), nb.end());
            return nb;
        };
        int u1 = verts[0];
        size_t deg1 = 0;
        for (int u : verts) {
            const size_t d = neighbours(u).size();
            if (d >= deg1) deg1 = d, u1 = u;
        }
        const std::vector<int> n1 = neighbours(u1);
        int best_u2 = -1, best_cov = -1;
        std::vector<int> best_vs;
        auto has = [](const std::vector<int>& sorted, int x) { return std::binary_search(sorted.begin(), sorted.end(), x); };
        for (size_t c = 0; c <= verts.size(); ++c) {
            const int u2 = c < verts.size() ? verts[c] : -1;
            if (u2 == u1) continue;
            const std::vector<int> n2 = u2 >= 0 ? neighbours(u2) : std::vector<int>();
            std::vector<int> cand = n1;
            cand.insert(cand.end(), n2.begin(), n2.end());
            std::sort(cand.begin(), cand.end());
            cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
            std::stable_sort(cand.begin(), cand.end(), [&](int x, int y) {
                return (has(n1, x) + has(n2, x)) > (has(n1, y) + has(n2, y));
            });
            if (cand.size() > 8) cand.resize(8);
            // terms covered (an edge {u1, u2} reachable both ways counts once)
            int cov = 0;
            for (const auto& e : left) {
                bool hit = false;
                for (int v : cand)
                    if ((e[1] == u1 && e[2] == v) || (e[2] == u1 && e[1] == v) || (u2 >= 0 && ((e[1] == u2 && e[2] == v) || (e[2] == u2 && e[1] == v))))
                        hit = true;
                cov += hit;
            }
            if (cov > best_cov) best_cov = cov, best_u2 = u2, best_vs = cand;
        }
        MbOp op;
        op.u[0] = u1;
        op.u[1] = best_u2;
        for (int j = 0; j < 8; ++j) {
            op.v[j] = j < (int)best_vs.size() ? best_vs[j] : -1;
            op.dst[0][j] = op.dst[1][j] = -1;
        }
        for (int r = 0; r < 2; ++r)
            for (int j = 0; j < 8; ++j) {
                if (op.u[r] < 0 || op.v[j] < 0) continue;
                for (size_t i = 0; i < left.size(); ++i) {
                    const auto& e = left[i];
                    if ((e[1] == op.u[r] && e[2] == op.v[j]) || (e[2] == op.u[r] && e[1] == op.v[j])) {
                        op.dst[r][j] = e[0];
                        left.erase(left.begin() + i);
                        break;
                    }
                }
            }
        ops->push_back(op);
    }
}

// Per-pair warp tables of gram_mb_kernel (layout: mma_common.cuh).  false: a virtual term has three factors, or some
// pair needs more ops / copy slots per warp than the kernel provides.
static bool build_mb_tables(MomentMap* M, int n, int n_t) {
    constexpr int kPitch = kBT + 4, kStage = kKS * kPitch;
    const int nblocks = M->nrb + M->ncb;
    const int row_terms = M->nrows + n_t, col_terms = n_t + M->ncols;
    auto factor_id = [&](uint16_t code) {
        if (code == kNoFactor) return 0;
        return (code & kFactorIsY) ? n + 1 + (int)(code & 0x7FFF) : (int)code + 1;
    };
    std::vector<std::vector<MbOp>> ops(nblocks);
    std::vector<std::vector<std::pair<int, int>>> singles(nblocks);  // (column, factor): values that are copied
    M->block_ops.assign(nblocks, 0);
    for (int b = 0; b < nblocks; ++b) {
        std::vector<std::array<int, 3>> terms;
        for (int c = 0; c < kBT; ++c) {
            const int idx = b * kBT + c;
            const bool valid = b < M->nrb ? idx < row_terms : idx - M->nrb * kBT < col_terms;
            if (!valid) continue;
            const uint16_t* f = &M->tab[(size_t)idx * kMaxFactors];
            if (f[2] != kNoFactor) return false;  // three factors: not a single product
            const int f0 = factor_id(f[0]), f1 = factor_id(f[1]);
            if (f0 == 0 || f1 == 0) singles[b].push_back({c, f0 == 0 ? f1 : f0});
            else terms.push_back({c, f0, f1});
        }
        cover_block(terms, &ops[b]);
        // a block of targets has more single values than the copy slots of a warp take: the rest is multiplied by
        // the constant, eight at a time
        const size_t keep = (size_t)(8 * kMbCopySlots * 32 / kKS) / 2;
        while (singles[b].size() > keep) {
            MbOp op;
            op.u[0] = 0;
            op.u[1] = -1;
            for (int j = 0; j < 8; ++j) {
                op.v[j] = op.dst[0][j] = op.dst[1][j] = -1;
                if (singles[b].size() > keep) {
                    op.v[j] = singles[b].back().second;
                    op.dst[0][j] = singles[b].back().first;
                    singles[b].pop_back();
                }
            }
            ops[b].push_back(op);
        }
        M->block_ops[b] = (int)ops[b].size();
    }
    auto src_off = [&](int f, int s) -> uint32_t {  // byte offset inside a raw-sample slot
        if (f < 0) return kVZero * 8;
        if (f == 0) return (uint32_t)(kVOnes + s) * 8;
        if (f <= n) return (uint32_t)(kVX + s * mb_pitch(n) + (f - 1)) * 8;
        return (uint32_t)(kVX + kKS * mb_pitch(n) + s * n_t + (f - n - 1)) * 8;
    };
    const int npairs = (int)M->pairs.size();
    M->optab.assign((size_t)npairs * kMbPairWords, 0u);
    M->opcnt.assign((size_t)npairs * 8, 0);
    for (int p = 0; p < npairs; ++p) {
        struct Inst {
            int tile, sg;
            const MbOp* op;
        };
        std::vector<Inst> inst;
        std::vector<uint32_t> copies;  // source | destination << 16
        const int blk[2] = {M->pairs[p].first, M->pairs[p].second};
        const int used = M->rows_used[p];
        for (int sg = 0; sg < kKS / 4; ++sg)
            for (int tile = 0; tile < 2; ++tile)
                for (const MbOp& op : ops[blk[tile]]) {
                    bool live = false;
                    for (int r = 0; r < 2; ++r)
                        for (int j = 0; j < 8; ++j) live |= op.dst[r][j] >= 0 && (tile == 1 || op.dst[r][j] < used);
                    if (live) inst.push_back({tile, sg, &op});
                }
        for (int tile = 0; tile < 2; ++tile)
            for (int s = 0; s < kKS; ++s)
                for (const auto& cf : singles[blk[tile]])
                    if (tile == 1 || cf.first < used)
                        copies.push_back(src_off(cf.second, s) | ((uint32_t)(tile * kStage + s * kPitch + cf.first) * 8) << 16);
        const int ncslots = ((int)copies.size() + 31) / 32;
        if ((int)inst.size() > 8 * kMbOpSlots || ncslots > 8 * kMbCopySlots) {
            if (getenv("DDB_GRAM_DEBUG"))
                fprintf(stderr, "[ddb] tensor-pipe builder: pair %d (blocks %d, %d) needs %d ops and %d copy slots per stage\n", p,
                        blk[0], blk[1], (int)inst.size(), ncslots);
            return false;
        }
        uint32_t* tab = &M->optab[(size_t)p * kMbPairWords];
        for (int w = 0; w < 8; ++w) {
            uint32_t* wt = tab + (size_t)w * kMbWarpWords;
            for (int j = 0; j < kMbOpSlots; ++j) {
                const int idx = j * 8 + w;           // instances are dealt to the warps round-robin
                const int kk = j % 4, slot = j / 4;  // a warp's first four are slot 0 of k-steps 0..3, the next four slot 1
                const Inst* in = idx < (int)inst.size() ? &inst[idx] : nullptr;
                if (in && slot == 2) M->opcnt[(size_t)p * 8 + w] |= (uint8_t)(1u << kk);
                for (int lane = 0; lane < 32; ++lane) {
                    // lane (g, t) holds A[g][t] = v_g(sample t), B[t][g] with column g = (sample g / 2, second factor
                    // g % 2) one-hot in t, and D[g][2t], D[g][2t + 1] = v_g u_0, v_g u_1 at sample t: the stores of a
                    // run of consecutive columns hit the banks like the fragment loads do (pitch = 4 mod 16)
                    const int g = lane >> 2, t = lane & 3;
                    const int sl = g >> 1, r = g & 1;
                    uint32_t offA = kVZero * 8, offB = kVZero * 8, d0, d1;
                    const int tile = in ? in->tile : 0, s0 = in ? in->sg * 4 : 0;
                    d0 = d1 = (uint32_t)(tile * kStage + (s0 + t) * kPitch + kBT + (g & 3)) * 8;  // pad columns: never read
                    if (in) {
                        const MbOp& op = *in->op;
                        offA = src_off(op.v[g], s0 + t);
                        if (t == sl) offB = src_off(op.u[r], s0 + sl);
                        const int c0 = op.dst[0][g], c1 = op.dst[1][g];
                        if (c0 >= 0 && (tile == 1 || c0 < used)) d0 = (uint32_t)(tile * kStage + (s0 + t) * kPitch + c0) * 8;
                        if (c1 >= 0 && (tile == 1 || c1 < used)) d1 = (uint32_t)(tile * kStage + (s0 + t) * kPitch + c1) * 8;
                    }
                    uint32_t* e = wt + ((kk * 3 + slot) * 32 + lane) * 2;
                    e[0] = offA | (offB << 16);
                    e[1] = d0 | (d1 << 16);
                }
            }
            int mine = 0;
            for (int j = 0; j < kMbCopySlots; ++j) {
                const int idx = j * 8 + w;
                if (idx < ncslots) mine = j + 1;
                for (int lane = 0; lane < 32; ++lane) {
                    const size_t ci = (size_t)idx * 32 + lane;
                    const uint32_t dump = (uint32_t)(kBT + (lane & 3)) * 8;
                    wt[kMbOpSlots * 64 + j * 32 + lane] = (idx < ncslots && ci < copies.size()) ? copies[ci] : (kVZero * 8) | (dump << 16);
                }
            }
            M->opcnt[(size_t)p * 8 + w] |= (uint8_t)(mine << 4);
        }
    }
    return true;
}

static void analyse_library(const uint8_t* exps, int n, int k, int n_t, int D, MomentMap* M) {
    M->worth = false;
    if (D < 1 || D > kMaxFactors || k > kMaxMomentTerms) return;
    // sorted factor codes of every term, right-aligned in D entries (code = state index + 1, 0 = constant)
    std::vector<Key> term(k);
    for (int j = 0; j < k; ++j) {
        Key key{};
        int d = 0;
        uint16_t f[kMaxFactors];
        for (int i = 0; i < n; ++i)
            for (int e = 0; e < exps[(size_t)i + (size_t)n * j]; ++e) f[d++] = (uint16_t)(i + 1);
        for (int i = 0; i < d; ++i) key[D - d + i] = f[i];
        term[j] = key;
    }
    std::unordered_map<Key, int, KeyHash> row_id, col_id;
    std::vector<Key> rows, cols;
    auto get = [](std::unordered_map<Key, int, KeyHash>& map, std::vector<Key>& list, const Key& key) {
        auto it = map.find(key);
        if (it != map.end()) return it->second;
        const int id = (int)list.size();
        map.emplace(key, id);
        list.push_back(key);
        return id;
    };
    // every library term is a row (R = Theta Y' pairs it with the target columns)
    for (int j = 0; j < k; ++j) get(row_id, rows, term[j]);
    // (row id, column id) of every lower-triangular Gram entry
    const size_t ntri = (size_t)k * (k + 1) / 2;
    std::vector<int32_t> er(ntri), ec(ntri);
    // A moment may be cut in several ways; any (P, Q) whose product is the moment will do.  In order:
    //   1. the D smallest / D largest of the padded factor list (densest staircase).  If over the whole library
    //      this cut needs only a few monomials that are not library terms (e.g. the constant of a library without
    //      one), they become virtual terms and cut 1 is used throughout.  Otherwise, for an entry whose cut 1
    //      leaves the library:
    //   2. the lower / upper half of the actual factors if both are library terms (mixed-degree libraries);
    //   3. the entry's own two terms.
    const std::unordered_map<Key, int, KeyHash> lib = row_id;
    auto in_lib = [&](const Key& key) { return lib.find(key) != lib.end(); };
    // cut `which` (1 or 2) of entry (t, u); returns whether both parts are library terms
    auto cut = [&](int t, int u, int which, Key& P, Key& Q) {
        uint16_t mg[2 * kMaxFactors];
        std::merge(term[t].begin(), term[t].begin() + D, term[u].begin(), term[u].begin() + D, mg);
        P = Key{};
        Q = Key{};
        if (which == 1) {
            for (int i = 0; i < D; ++i) {
                P[i] = mg[i];
                Q[i] = mg[D + i];
            }
        } else {
            int z = 0;
            while (z < 2 * D && mg[z] == 0) ++z;
            const int L = 2 * D - z, lp = L / 2, lq = L - lp;
            for (int i = 0; i < lp; ++i) P[D - lp + i] = mg[z + i];
            for (int i = 0; i < lq; ++i) Q[D - lq + i] = mg[z + lp + i];
        }
        return in_lib(P) && in_lib(Q);
    };
    std::vector<uint8_t> choice(ntri, 0);  // 0 = neither cut stays inside the library
    std::unordered_map<Key, int, KeyHash> foreign;
    size_t e = 0;
    for (int t = 0; t < k; ++t)
        for (int u = 0; u <= t; ++u, ++e) {
            Key P, Q;
            if (cut(t, u, 1, P, Q)) {
                choice[e] = 1;
            } else {
                if (!in_lib(P)) foreign.emplace(P, 0);
                if (!in_lib(Q)) foreign.emplace(Q, 0);
                if (cut(t, u, 2, P, Q)) choice[e] = 2;
            }
        }
    const bool allow_foreign = foreign.size() <= (size_t)(32 + (k + n_t) / 16);
    e = 0;
    for (int t = 0; t < k; ++t)
        for (int u = 0; u <= t; ++u, ++e) {
            Key P, Q;
            if (allow_foreign || choice[e] == 1) {
                cut(t, u, 1, P, Q);
            } else if (choice[e] == 2) {
                cut(t, u, 2, P, Q);
            } else {
                P = term[t];
                Q = term[u];
            }
            er[e] = get(row_id, rows, P);
            ec[e] = get(col_id, cols, Q);
        }
    // positions: rows by increasing max (then lexicographic), targets last; columns: targets first, then the
    // full-degree ones by decreasing min, then the lower-degree ones by decreasing min (then lexicographic)
    const int nrows = (int)rows.size(), ncols = (int)cols.size();
    std::vector<int> rorder(nrows), corder(ncols), rpos(nrows), cpos(ncols);
    for (int i = 0; i < nrows; ++i) rorder[i] = i;
    for (int i = 0; i < ncols; ++i) corder[i] = i;
    std::sort(rorder.begin(), rorder.end(), [&](int x, int y) {
        if (rows[x][D - 1] != rows[y][D - 1]) return rows[x][D - 1] < rows[y][D - 1];
        return rows[x] < rows[y];
    });
    auto min_factor = [&](const Key& key) {  // smallest actual factor (0 for the constant)
        for (int i = 0; i < D; ++i)
            if (key[i]) return (int)key[i];
        return 0;
    };
    std::sort(corder.begin(), corder.end(), [&](int x, int y) {
        if (cols[x][0] != cols[y][0]) return cols[x][0] > cols[y][0];  // full-degree columns first
        const int mx = min_factor(cols[x]), my = min_factor(cols[y]);
        if (mx != my) return mx > my;
        return cols[x] < cols[y];
    });
    for (int i = 0; i < nrows; ++i) rpos[rorder[i]] = i;
    for (int i = 0; i < ncols; ++i) cpos[corder[i]] = n_t + i;
    const int nr_total = nrows + n_t, nc_total = n_t + ncols;
    const int nrb = (nr_total + kBT - 1) / kBT, ncb = (nc_total + kBT - 1) / kBT;
    // needed block pairs
    // bits 0..3: a needed entry in rows [32 q, 32 q + 32) of the block pair, bits 4..7: the same for columns
    std::vector<int> pair_of((size_t)nrb * ncb, 0);
    auto mark = [&](int r, int c) {
        pair_of[(size_t)(r / kBT) * ncb + c / kBT] |= (1 << ((r % kBT) / 32)) | (16 << ((c % kBT) / 32));
    };
    for (size_t i = 0; i < ntri; ++i) mark(rpos[er[i]], cpos[ec[i]]);
    for (int t = 0; t < k; ++t)
        for (int j = 0; j < n_t; ++j) mark(rpos[row_id[term[t]]], j);
    for (int j = 0; j < n_t; ++j) mark(nrows + j, j);
    // whole pairs first, then the pairs at the edge of the staircase that only need their first 64, then those
    // that only need their first 32 rows -- or columns: those run transposed (column block as the kernel's row
    // operand) -- as the kernel's 64 x 128 / 32 x 128 patches; pair_of becomes the pair index
    M->pairs.clear();
    M->rows_used.clear();
    std::vector<uint8_t> transposed;
    auto extent = [](int bits) { return (bits & 8) ? 128 : (bits & 4) ? 96 : (bits & 2) ? 64 : 32; };
    for (int pass = 0; pass < 3; ++pass)
        for (int I = 0; I < nrb; ++I)
            for (int J = 0; J < ncb; ++J) {
                int& cell = pair_of[(size_t)I * ncb + J];
                if (cell <= 0) continue;
                const int re = extent(cell & 15), ce = extent((cell >> 4) & 15);
                // (a 96-row class was built and measured in round 2: the 96 x 128 patch costs 0.92 of a whole pair for
                // 0.75 of its flops, and the seventh patch shape alone made gram_kernel 2.8 % slower -- not kept)
                const int rcls = re <= 32 ? 32 : re <= 64 ? 64 : 128, ccls = ce <= 32 ? 32 : ce <= 64 ? 64 : 128;
                const int used = std::min(rcls, ccls);
                if (used != (pass == 0 ? 128 : pass == 1 ? 64 : 32)) continue;
                const bool tr = ccls < rcls;
                cell = -(int)M->pairs.size() - 1;  // negative while the passes run: index = -cell - 1
                if (tr) M->pairs.push_back({nrb + J, I});
                else M->pairs.push_back({I, nrb + J});
                M->rows_used.push_back((uint8_t)used);
                transposed.push_back(tr);
            }
    for (int& cell : pair_of) cell = cell < 0 ? -cell - 1 : -1;
    const int npairs = (int)M->pairs.size();
    M->nrows = nrows;
    M->ncols = ncols;
    M->nrb = nrb;
    M->ncb = ncb;
    M->nvirt = nrb * kBT + nc_total;
    // the plain schedule: nblk (nblk - 1) / 2 full pairs + nblk diagonal pairs at 0.6
    const int nblk = (k + n_t + kBT - 1) / kBT;
    const double plain = 0.5 * nblk * (nblk - 1) + 0.6 * nblk;
    if (getenv("DDB_GRAM_DEBUG"))
        fprintf(stderr, "[ddb] moment schedule: k=%d n_t=%d D=%d rows=%d cols=%d blocks %d x %d pairs=%d: %d whole, %d half, %d quarter (plain %.1f)\n",
                k, n_t, D, nrows, ncols, nrb, ncb, npairs, (int)std::count(M->rows_used.begin(), M->rows_used.end(), (uint8_t)128),
                (int)std::count(M->rows_used.begin(), M->rows_used.end(), (uint8_t)64),
                (int)std::count(M->rows_used.begin(), M->rows_used.end(), (uint8_t)32), plain);
    double cost = 0.0;
    for (uint8_t h : M->rows_used) cost += pair_cost(h);
    if (npairs >= (1 << 17) || cost > 0.9 * plain) return;

    // gather map of the packed [G | R | yy]
    const size_t count = (size_t)k * k + (size_t)k * n_t + n_t;
    std::vector<int32_t>& src = M->src;
    src.assign(count, 0);
    auto src_of = [&](int r, int c) {
        const int p = pair_of[(size_t)(r / kBT) * ncb + c / kBT];
        const int off = transposed[p] ? (c % kBT) * kBT + (r % kBT) : (r % kBT) * kBT + (c % kBT);
        return (int32_t)((p << 14) | off);
    };
    e = 0;
    for (int t = 0; t < k; ++t)
        for (int u = 0; u <= t; ++u, ++e) {
            const int32_t s = src_of(rpos[er[e]], cpos[ec[e]]);
            src[(size_t)t + (size_t)k * u] = s;
            src[(size_t)u + (size_t)k * t] = s;
        }
    for (int j = 0; j < n_t; ++j) {
        for (int t = 0; t < k; ++t) src[(size_t)k * k + (size_t)t + (size_t)k * j] = src_of(rpos[row_id[term[t]]], j);
        src[(size_t)k * k + (size_t)k * n_t + j] = src_of(nrows + j, j);
    }
    // virtual factor table: [rows | targets | padding to a block] [targets | columns]
    const int tab_rows = (nrb + ncb) * kBT;
    std::vector<uint16_t>& tab = M->tab;
    tab.assign((size_t)tab_rows * kMaxFactors, kNoFactor);
    auto put = [&](int row, const Key& key) {
        int d = 0;
        for (int i = 0; i < D; ++i)
            if (key[i]) tab[(size_t)row * kMaxFactors + d++] = (uint16_t)(key[i] - 1);
    };
    for (int i = 0; i < nrows; ++i) put(rpos[i], rows[i]);
    for (int j = 0; j < n_t; ++j) {
        tab[(size_t)(nrows + j) * kMaxFactors] = (uint16_t)(kFactorIsY | j);
        tab[(size_t)(nrb * kBT + j) * kMaxFactors] = (uint16_t)(kFactorIsY | j);
    }
    for (int i = 0; i < ncols; ++i) put(nrb * kBT + cpos[i], cols[i]);
    M->worth = true;
    // products on the tensor pipe: virtual terms of at most two factors
    if (D > 2 || !build_mb_tables(M, n, n_t)) {
        M->optab.clear();
        M->opcnt.clear();
    }
}

static int install_library(ddb_ctx* ctx, MomentState* M) {
    M->eligible = false;
    M->lib_version = ctx->lib_version;
    M->k = ctx->k;
    M->n_t = ctx->n_t;
    MomentMap map;
    analyse_library(ctx->exps.data(), ctx->n, ctx->k, ctx->n_t, ctx->max_degree, &map);
    if (!map.worth) return DDB_OK;
    M->pairs = map.pairs;
    M->npairs = (int)map.pairs.size();
    M->rows_used = map.rows_used;
    M->nrb = map.nrb;
    M->ncb = map.ncb;
    M->nvirt = map.nvirt;
    DDB_TRY(M->factors.reserve(map.tab.size() * sizeof(uint16_t)));
    DDB_TRY(M->src.reserve(map.src.size() * sizeof(int32_t)));
    DDB_CUDA_TRY(cudaMemcpyAsync(M->factors.p, map.tab.data(), map.tab.size() * sizeof(uint16_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
    DDB_CUDA_TRY(cudaMemcpyAsync(M->src.p, map.src.data(), map.src.size() * sizeof(int32_t), cudaMemcpyHostToDevice,
                                 ctx->stream));
    M->mb = !map.optab.empty() && gram_mb_fits(ctx->n, ctx->n_t);
    if (M->mb) {
        DDB_TRY(M->optab.reserve(map.optab.size() * sizeof(uint32_t)));
        DDB_TRY(M->opcnt.reserve(map.opcnt.size()));
        DDB_CUDA_TRY(cudaMemcpyAsync(M->optab.p, map.optab.data(), map.optab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                     ctx->stream));
        DDB_CUDA_TRY(cudaMemcpyAsync(M->opcnt.p, map.opcnt.data(), map.opcnt.size(), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (getenv("DDB_GRAM_DEBUG")) {
        fprintf(stderr, "[ddb] moment schedule: tensor-pipe builder %s; ops per block:", M->mb ? "on" : "off");
        for (int c : map.block_ops) fprintf(stderr, " %d", c);
        fprintf(stderr, "\n");
    }
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));  // `map` is a stack-lifetime host buffer
    M->eligible = true;
    for (auto& pl : M->plans) pl.m = -1;
    return DDB_OK;
}

// Host-only view of the analysis for ddb_moment_plan_describe (api.cu).
int64_t moment_plan_describe(const uint8_t* exps, int n, int k, int n_t, int32_t* src, int64_t src_cap, uint16_t* tab,
                             int64_t tab_cap, int32_t* pairs, int64_t pairs_cap, int64_t* info) {
    int D = 0;
    for (int j = 0; j < k; ++j) {
        int d = 0;
        for (int i = 0; i < n; ++i) d += exps[(size_t)i + (size_t)n * j];
        D = std::max(D, d);
    }
    MomentMap map;
    analyse_library(exps, n, k, n_t, D, &map);
    if (info) {
        info[0] = map.worth, info[1] = map.nrows, info[2] = map.ncols, info[3] = map.nrb, info[4] = map.ncb;
        info[5] = (int64_t)map.pairs.size(), info[6] = map.nvirt, info[7] = D;
    }
    if (!map.worth) return 0;
    if (src)
        for (int64_t i = 0; i < std::min<int64_t>(src_cap, (int64_t)map.src.size()); ++i) src[i] = map.src[i];
    if (tab)
        for (int64_t i = 0; i < std::min<int64_t>(tab_cap, (int64_t)map.tab.size()); ++i) tab[i] = map.tab[i];
    if (pairs)
        for (int64_t i = 0; i < std::min<int64_t>(pairs_cap, (int64_t)map.pairs.size()); ++i) {
            pairs[3 * i] = map.pairs[i].first;
            pairs[3 * i + 1] = map.pairs[i].second;
            pairs[3 * i + 2] = map.rows_used[i];
        }
    return (int64_t)map.src.size();
}

// Host-only view of the tensor-pipe builder tables for ddb_moment_builder_describe (api.cu).
int64_t moment_builder_describe(const uint8_t* exps, int n, int k, int n_t, uint32_t* optab, int64_t optab_cap,
                                uint8_t* opcnt, int64_t opcnt_cap, int32_t* block_ops, int64_t block_cap) {
    int D = 0;
    for (int j = 0; j < k; ++j) {
        int d = 0;
        for (int i = 0; i < n; ++i) d += exps[(size_t)i + (size_t)n * j];
        D = std::max(D, d);
    }
    MomentMap map;
    analyse_library(exps, n, k, n_t, D, &map);
    if (!map.worth || map.optab.empty() || !gram_mb_fits(n, n_t)) return 0;
    if (optab)
        for (int64_t i = 0; i < std::min<int64_t>(optab_cap, (int64_t)map.optab.size()); ++i) optab[i] = map.optab[i];
    if (opcnt)
        for (int64_t i = 0; i < std::min<int64_t>(opcnt_cap, (int64_t)map.opcnt.size()); ++i) opcnt[i] = map.opcnt[i];
    if (block_ops)
        for (int64_t i = 0; i < std::min<int64_t>(block_cap, (int64_t)map.block_ops.size()); ++i) block_ops[i] = map.block_ops[i];
    return (int64_t)map.pairs.size();
}

// Lock-step waves.  A wave gives every one of its block pairs a number of equal sample ranges, one CTA per
// (pair, range), at most `nctas` CTAs in all; the CTAs of a range walk the same samples at the same time, so a
// wave reads the X / Y stream from HBM about once per cost class.  Waves follow each other inside one launch (a
// CTA's segment list is its share of every wave).  Two kinds of wave, chosen by dynamic programming over "how
// many of the most expensive pairs go first" (the pairs are sorted whole, half, quarter):
//   * a homogeneous wave of np = min(rest of the class, nctas / nr) equal-cost pairs over nr ranges each, any nr;
//   * a FINAL wave of all remaining pairs, pair p over r_p = ceil(cost_p / T) ranges with the smallest T that
//     fits -- CTAs of cheaper pairs get longer ranges so that all take about the time T.
struct WavePlan {
    std::vector<std::vector<GramSegment>> per_cta;
    std::vector<int32_t> psb;  // pair -> first slot (slots of a pair are contiguous), npairs + 1
    int nslots = 0, nwaves = 0;
    double time = 0.0;  // pair streams per CTA (1.0 = one whole pair over all samples)
};

static void plan_waves(const std::vector<std::pair<int, int>>& pairs, const std::vector<uint8_t>& rows_used, int64_t ns,
                       int nctas, WavePlan* W) {
    const int npairs = (int)pairs.size();
    auto cls_of = [&](int p) { return pair_class(rows_used[p]); };
    auto cost_of = [&](int p) { return pair_cost(rows_used[p]); };
    const int64_t max_r = std::max<int64_t>(1, ns);
    double ideal = 0.0;
    for (int p = 0; p < npairs; ++p) ideal += cost_of(p) / nctas;
    const double kWavePenalty = kWavePenaltyFrac * ideal;
    // final wave from `first`: ranges per pair and its time; false if even one range per pair does not fit
    auto final_wave = [&](int first, std::vector<int>* ranges, double* T) {
        if (npairs - first > nctas) return false;
        auto count = [&](double t, std::vector<int>* out) {
            int64_t total = 0;
            for (int p = first; p < npairs; ++p) {
                const int64_t r = std::min<int64_t>(max_r, std::max<int64_t>(1, (int64_t)std::ceil(cost_of(p) / t - 1e-12)));
                if (out) (*out)[p - first] = (int)r;
                total += r;
            }
            return total;
        };
        double lo = 0.0, hi = 1.0;  // count(hi = 1) = one range per pair: fits
        for (int it = 0; it < 60; ++it) {
            const double mid = 0.5 * (lo + hi);
            if (count(mid, nullptr) <= nctas) hi = mid;
            else lo = mid;
        }
        ranges->assign(npairs - first, 1);
        count(hi, ranges);
        *T = 0.0;
        for (int p = first; p < npairs; ++p) *T = std::max(*T, cost_of(p) / (*ranges)[p - first]);
        return true;
    };
    // homogeneous wave from `first` with nr ranges per pair: how many pairs it takes (0 = not possible)
    auto class_end = [&](int first) {
        int e = first;
        while (e < npairs && rows_used[e] == rows_used[first]) ++e;
        return e;
    };
    auto homogeneous = [&](int first, int nr) { return std::min(class_end(first) - first, nctas / nr); };
    // best[first] = least time for the pairs [first, npairs) (every wave also pays kWavePenalty);
    // choice[first] = 0: final mixed wave, nr > 0: homogeneous wave with nr ranges per pair
    std::vector<double> best(npairs + 1, 0.0);
    std::vector<int> choice(npairs + 1, 0);
    const int nr_max = (int)std::min<int64_t>(nctas, max_r);
    for (int first = npairs - 1; first >= 0; --first) {
        best[first] = 1e300;
        for (int nr = 1; nr <= nr_max; ++nr) {
            const int np = homogeneous(first, nr);
            if (np < 1) break;
            const double t = cost_of(first) / nr + kWavePenalty + best[first + np];
            if (t < best[first] * (1.0 - 1e-12)) {
                best[first] = t;
                choice[first] = nr;
            }
        }
        std::vector<int> ranges;
        double T;
        if (final_wave(first, &ranges, &T) && T + kWavePenalty <= best[first] * (1.0 + 1e-9)) {  // ties: fewer waves
            best[first] = T + kWavePenalty;
            choice[first] = 0;
        }
    }
    W->per_cta.assign(nctas, {});
    W->psb.assign(npairs + 1, 0);
    W->nslots = 0;
    W->nwaves = 0;
    std::vector<double> used(nctas, 0.0);  // pair streams a CTA has been given so far
    auto push = [&](int p, int cta, int64_t b, int64_t e) {
        GramSegment sg;
        sg.bi = pairs[p].first;
        sg.bj = pairs[p].second;
        sg.slot = W->nslots++;
        sg.pad = cls_of(p) | (p << 8);  // bits 0..1: 1 = 64 x 128 patch, 2 = 32 x 128 patch; bits 8..: pair index
        sg.stage_begin = b;
        sg.stage_end = e;
        W->per_cta[cta].push_back(sg);
        used[cta] += cost_of(p) * (double)(e - b) / (double)max_r;
    };
    int first = 0;
    while (first < npairs) {
        ++W->nwaves;
        if (choice[first] != 0) {
            const int nr = choice[first], np = homogeneous(first, nr);
            int cta = 0;
            for (int p = first; p < first + np; ++p) {
                W->psb[p] = W->nslots;
                for (int r = 0; r < nr; ++r) push(p, cta++, ns * r / nr, ns * (r + 1) / nr);
            }
            first += np;
            continue;
        }
        // Final wave.  The DP sized it as "pair p over ceil(cost_p / T) equal ranges"; it is dealt out so that every CTA
        // finishes at a common time (the water level below), the CTAs the earlier waves left idle included: no
        // quantisation by class, at the price of a few CTAs that change pair inside the wave (a flush and a pipeline
        // fill, microseconds against a wave of tens of milliseconds).
        if (getenv("DDB_MOMENT_STREAM") && atoi(getenv("DDB_MOMENT_STREAM")) == 0) {  // as sized: equal ranges per pair
            std::vector<int> ranges;
            double T;
            final_wave(first, &ranges, &T);
            int cta = 0;
            for (int p = first; p < npairs; ++p) {
                W->psb[p] = W->nslots;
                for (int r = 0; r < ranges[p - first]; ++r) push(p, cta++, ns * r / ranges[p - first], ns * (r + 1) / ranges[p - first]);
            }
            first = npairs;
            continue;
        }
        double work = 0.0;
        for (int p = first; p < npairs; ++p) work += cost_of(p);
        std::vector<double> sorted_used = used;
        std::sort(sorted_used.begin(), sorted_used.end());
        double level = sorted_used[0] + work, prefix = 0.0;  // water level: sum_c max(0, level - used_c) = work
        for (int c = 0; c < nctas; ++c) {
            prefix += sorted_used[c];
            const double lv = (work + prefix) / (c + 1);
            if (c + 1 == nctas || lv <= sorted_used[c + 1]) {
                level = lv;
                break;
            }
        }
        // Two phases.  (1) Aligned pieces: a CTA with the standard room B (every CTA that took part in all earlier waves)
        // takes L_p = B / (cost per stage) stages of ONE pair, at a multiple of L_p -- the pairs of a cost class then walk
        // the same samples at the same time and share their X / Y tiles through L2, as in the homogeneous waves.
        // (2) What is left of every pair (less than one piece) is dealt out as a stream over the remaining CTAs, each up
        // to the level; a piece shorter than kMinPiece stages is not worth a segment.  Pieces are whole stages, so with
        // few stages per pair the last CTA would collect what the roundings left over: the level is raised (bisection
        // on a dry run) until the stream ends within it.
        const int64_t kMinPiece = std::max<int64_t>(1, std::min<int64_t>(32, ns / 64));
        struct Piece {
            int p, cta;
            int64_t b, e;
        };
        double std_used = 0.0;  // the most common load so far
        {
            int best_n = 0;
            for (int c = 0; c < nctas; ++c) {
                int cnt = 0;
                for (int d = 0; d < nctas; ++d) cnt += std::fabs(used[d] - used[c]) < 1e-12;
                if (cnt > best_n) best_n = cnt, std_used = used[c];
            }
        }
        auto deal = [&](double lvl, std::vector<Piece>* out) {  // returns the finishing time of the last CTA of the stream
            std::vector<double> u = used;
            std::vector<char> taken(nctas, 0);
            std::vector<int64_t> rest(npairs - first, 0);  // first stage of the pair that phase 1 left
            const double room_std = lvl - std_used;
            int next_std = 0;
            for (int p = first; p < npairs; ++p) {
                const double per_stage = cost_of(p) / (double)max_r;
                const int64_t L = (int64_t)std::floor(room_std / per_stage + 1e-9);
                int64_t st = 0;
                while (L >= std::max<int64_t>(kMinPiece, 1) && ns - st >= L) {
                    while (next_std < nctas && (taken[next_std] || std::fabs(used[next_std] - std_used) >= 1e-12)) ++next_std;
                    if (next_std >= nctas) break;
                    taken[next_std] = 1;
                    u[next_std] += per_stage * (double)L;
                    if (out) out->push_back({p, next_std, st, st + L});
                    st += L;
                }
                rest[p - first] = st;
            }
            std::vector<int> free_ctas;
            for (int c = 0; c < nctas; ++c)
                if (!taken[c]) free_ctas.push_back(c);
            if (free_ctas.empty()) free_ctas.push_back((int)(std::min_element(u.begin(), u.end()) - u.begin()));
            size_t fi = 0;
            for (int p = first; p < npairs; ++p) {
                const double per_stage = cost_of(p) / (double)max_r;
                int64_t st = rest[p - first];
                while (st < ns) {
                    while (fi + 1 < free_ctas.size() && lvl - u[free_ctas[fi]] < per_stage * (double)std::min<int64_t>(kMinPiece, ns - st)) ++fi;
                    const int cta = free_ctas[fi];
                    int64_t take = ns - st;
                    if (fi + 1 < free_ctas.size()) {
                        take = std::min<int64_t>(take, (int64_t)std::floor((lvl - u[cta]) / per_stage + 0.5));
                        if (ns - st - take < kMinPiece && ns - st - take > 0) take = ns - st;  // no crumbs at a pair's end
                        take = std::max<int64_t>(take, 1);
                    }
                    if (out) out->push_back({p, cta, st, st + take});
                    u[cta] += per_stage * (double)take;
                    st += take;
                }
            }
            return u[free_ctas[fi]];
        };
        double max_stage = 0.0;
        for (int p = first; p < npairs; ++p) max_stage = std::max(max_stage, cost_of(p) / (double)max_r);
        double lo = level, hi = level + 2.0 * max_stage * (double)kMinPiece + 1e-12;
        while (deal(hi, nullptr) > hi + 0.5 * max_stage) hi += max_stage * (double)kMinPiece;  // (ends: one CTA can take everything)
        if (deal(lo, nullptr) > lo + 0.5 * max_stage) {
            for (int it = 0; it < 40; ++it) {
                const double mid = 0.5 * (lo + hi);
                if (deal(mid, nullptr) > mid + 0.5 * max_stage) lo = mid;
                else hi = mid;
            }
            level = hi;
        }
        std::vector<Piece> pieces;
        deal(level, &pieces);
        // slots of a pair: contiguous, in stage order
        std::vector<int> order(pieces.size()), slot_of(pieces.size());
        for (size_t i = 0; i < pieces.size(); ++i) order[i] = (int)i;
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
            return pieces[x].p != pieces[y].p ? pieces[x].p < pieces[y].p : pieces[x].b < pieces[y].b;
        });
        for (size_t i = 0; i < order.size(); ++i) {
            const Piece& pc = pieces[order[i]];
            if (i == 0 || pieces[order[i - 1]].p != pc.p) W->psb[pc.p] = W->nslots + (int)i;
            slot_of[order[i]] = W->nslots + (int)i;
        }
        for (size_t i = 0; i < pieces.size(); ++i) {
            const Piece& pc = pieces[i];
            push(pc.p, pc.cta, pc.b, pc.e);
            W->per_cta[pc.cta].back().slot = slot_of[i];
        }
        first = npairs;
    }
    W->psb[npairs] = W->nslots;
    W->time = *std::max_element(used.begin(), used.end());
}

static int build_wave_plan(ddb_ctx* ctx, MomentState* M, MomentState::DevPlan** out) {
    const int nctas = ctx->sm_count;
    MomentState::DevPlan* P = nullptr;
    for (auto& pl : M->plans)
        if (pl.m == ctx->m && pl.nctas == nctas) P = &pl;
    if (P) {
        P->stamp = ++M->clock;
        *out = P;
        return DDB_OK;
    }
    P = &M->plans[0];
    for (auto& pl : M->plans)
        if (pl.stamp < P->stamp) P = &pl;  // least recently used
    const int64_t ns = (ctx->m + kKS - 1) / kKS;
    WavePlan W;
    plan_waves(M->pairs, M->rows_used, ns, nctas, &W);
    const std::vector<std::vector<GramSegment>>& per_cta = W.per_cta;
    const std::vector<int32_t>& psb = W.psb;
    const int nslots = W.nslots;
    std::vector<GramSegment> segs;
    std::vector<int32_t> cta_seg_begin(nctas + 1, 0);
    for (int c = 0; c < nctas; ++c) {
        cta_seg_begin[c] = (int32_t)segs.size();
        segs.insert(segs.end(), per_cta[c].begin(), per_cta[c].end());
    }
    cta_seg_begin[nctas] = (int32_t)segs.size();
    P->m = -1;
    P->off_cta = segs.size() * sizeof(GramSegment);
    P->off_psb = P->off_cta + cta_seg_begin.size() * sizeof(int32_t);
    const size_t total = P->off_psb + psb.size() * sizeof(int32_t);
    std::vector<unsigned char> host(total);
    memcpy(host.data(), segs.data(), segs.size() * sizeof(GramSegment));
    memcpy(host.data() + P->off_cta, cta_seg_begin.data(), cta_seg_begin.size() * sizeof(int32_t));
    memcpy(host.data() + P->off_psb, psb.data(), psb.size() * sizeof(int32_t));
    DDB_TRY(P->buf.reserve(total));
    DDB_CUDA_TRY(cudaMemcpyAsync(P->buf.p, host.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    DDB_TRY(M->partials.reserve((size_t)nslots * kBT * kBT * sizeof(double)));  // grows only: valid for every kept plan
    P->nslots = nslots;
    P->m = ctx->m;
    P->nctas = nctas;
    P->stamp = ++M->clock;
    *out = P;
    return DDB_OK;
}

int gram_moment_run(ddb_ctx* ctx, double* dPacked, bool* handled) {
    *handled = false;
    const char* env = getenv("DDB_GRAM_MOMENT");
    if (env && atoi(env) == 0) return DDB_OK;
    const int k = ctx->k, n_t = ctx->n_t;
    if (k + n_t <= 2 * kBT) return DDB_OK;  // one or two blocks: nothing to skip
    if (!ctx->moment) ctx->moment = new MomentState();
    MomentState* M = ctx->moment;
    if (M->lib_version != ctx->lib_version || M->k != k || M->n_t != n_t) DDB_TRY(install_library(ctx, M));
    if (!M->eligible) return DDB_OK;
    MomentState::DevPlan* P = nullptr;
    DDB_TRY(build_wave_plan(ctx, M, &P));
    // the library's own factor table (the rss / residual passes evaluate the real terms)
    if (ctx->factors_nt != n_t) DDB_TRY(gram_ensure_factors(ctx, (k + n_t + kBT - 1) / kBT * kBT));
    if (!dPacked) {
        DDB_TRY(ctx->d_packed.reserve((size_t)ddb_gram_count(ctx) * sizeof(double)));
        dPacked = ctx->d_packed.as<double>();
    }
    unsigned char* base = static_cast<unsigned char*>(P->buf.p);
    GramArgs a;
    a.X = ctx->dX;
    a.Y = ctx->dY;
    a.n = ctx->n;
    a.n_t = n_t;
    a.m = ctx->m;
    a.kaug = M->nvirt;
    a.factors = M->factors.as<uint16_t>();
    a.segs = reinterpret_cast<const GramSegment*>(base);
    a.cta_seg_begin = reinterpret_cast<const int32_t*>(base + P->off_cta);
    a.partials = M->partials.as<double>();
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[0], ctx->stream));
    // opt-in (DDB_GRAM_MB=1): bit-identical, the tensor pipe runs 3.4 points busier without the FP64 multiplies, but at
    // the 11 ops per warp and stage the greedy cover needs at C3 the extra DMMAs cost more than the multiplies did
    // (150.9 against 139.7 ms per 2e6 samples, profiles/gram_mb_experiment_r2.txt)
    const char* mb_env = getenv("DDB_GRAM_MB");
    if (M->mb && mb_env && atoi(mb_env) == 1) {
        a.optab = M->optab.as<uint32_t>();
        a.opcnt = M->opcnt.as<uint8_t>();
        DDB_TRY(gram_launch_mb128(ctx, a, ctx->sm_count));
    } else {
        DDB_TRY(gram_launch_bt128(ctx, a, ctx->sm_count, true));
    }
    const int64_t count = ddb_gram_count(ctx);
    moment_gather_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(
        M->partials.as<double>(), reinterpret_cast<const int32_t*>(base + P->off_psb), M->src.as<int32_t>(), count,
        dPacked);
    DDB_CUDA_TRY(cudaGetLastError());
    DDB_CUDA_TRY(cudaEventRecord(ctx->ev[1], ctx->stream));
    DDB_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    DDB_CUDA_TRY(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]));
    ctx->timings.gram_ms = ms;
    ctx->timings.gram_launches = 2;
    ctx->timings.gram_schedule = 2;
    double live_rows = 0.0;
    for (uint8_t h : M->rows_used) live_rows += h;
    ctx->timings.gram_dmma_flops = 2.0 * kBT * live_rows;
    ctx->have_gram = (dPacked == ctx->d_packed.as<double>());
    ctx->imp_k = 0;
    *handled = true;
    return DDB_OK;
}

// Host-only view of the wave plan for ddb_moment_wave_describe (api.cu): 7 int64 per segment.
int64_t moment_wave_describe(const uint8_t* exps, int n, int k, int n_t, int64_t m, int nctas, int64_t* segments,
                             int64_t cap, double* info) {
    int D = 0;
    for (int j = 0; j < k; ++j) {
        int d = 0;
        for (int i = 0; i < n; ++i) d += exps[(size_t)i + (size_t)n * j];
        D = std::max(D, d);
    }
    MomentMap map;
    analyse_library(exps, n, k, n_t, D, &map);
    if (!map.worth) return 0;
    WavePlan W;
    plan_waves(map.pairs, map.rows_used, (m + kKS - 1) / kKS, nctas, &W);
    double cost = 0.0;
    for (uint8_t h : map.rows_used) cost += pair_cost(h);
    if (info) info[0] = W.time, info[1] = cost / nctas, info[2] = W.nwaves, info[3] = W.nslots;
    int64_t w = 0;
    for (int c = 0; c < nctas; ++c)
        for (const GramSegment& g : W.per_cta[c]) {
            if (segments && w < cap) {
                int64_t* o = segments + 7 * w;
                o[0] = c, o[1] = g.bi, o[2] = g.bj, o[3] = g.slot, o[4] = g.stage_begin, o[5] = g.stage_end;
                o[6] = (g.pad & 3) == 0 ? 128 : (g.pad & 3) == 1 ? 64 : 32;
            }
            ++w;
        }
    return w;
}

void moment_free(ddb_ctx* ctx) {
    if (ctx->moment) {
        ctx->moment->factors.release();
        ctx->moment->src.release();
        ctx->moment->optab.release();
        ctx->moment->opcnt.release();
        for (auto& pl : ctx->moment->plans) pl.buf.release();
        ctx->moment->partials.release();
        delete ctx->moment;
        ctx->moment = nullptr;
    }
}

}  // namespace ddb
