// Shared declarations of the sweep / rss / select stages.
#pragma once
#include "common.cuh"

namespace ddb {

constexpr int kEqRunning = 0;
constexpr int kEqDone = 1;
constexpr int kEqNeedBig = 2;

// State machine of one target along the threshold sweep (device resident).
struct EqState {
    int j;          // current threshold index
    int iter;       // iterations done at this threshold
    int counter;    // total iterations (solver.jl:96)
    int na;         // size of the active set
    int status;     // kEqRunning / kEqDone / kEqNeedBig
    int nruns;      // recorded runs (maximal step sequences with unchanged cache content)
    int last_cand;  // candidate id of the current run
    int dof0;       // dof of the zeroed best cache
    int flags;      // bit 0: a factorisation broke down (no fallback applied), bit 1: minimum-norm steps were taken
    int pad[3];     // [0]: the initial solve is pending (runs as a masked step), [1]: parked for a minimum-norm big step
};

struct SweepState {
    int k = 0, n_t = 0, L = 0, W = 0;
    int64_t m_total = 0;
    ddb_sweep_params prm = {};
    bool valid = false;
    int ncand = 0;        // candidates recorded by the last sweep (host copy)
    int64_t pool_used = 0;
    int cand_cap = 0, run_cap = 0;
    int64_t pool_cap = 0;
    bool have_rss = false;
    std::vector<double> lambdas;

    DevBuf As, Lfac, dinv, rs, rraw, x, xprev, aux1, aux2, znew, active, eq, lambdas_d;
    DevBuf cand_idx, cand_val, cand_off, cand_nnz, cand_dof, cand_eq, counters, pool_used_d;
    DevBuf run_cand, run_jfirst, run_j;
    DevBuf path_support, path_coef, path_iters;
    DevBuf big_ws, big_idx, big_na, big_list, big_fail, active_matrix, shared_fail;
    DevBuf pinv_ws;       // un-scaled masked block, eigenvalues and two vectors of a minimum-norm big step
    DevBuf big_loc;   // sharded big steps: [na of my slots | fail of my slots | global slot of my slots] (3 x n_t ints)
    DevBuf dense_ids, dense_work;  // rss pass of dense candidates (rss_dense.cu): candidate ids; coefficient matrix + partial sums
    DevBuf big_xbuf;  // sharded big steps: per parked target [solution (k) | failure flag], summed over the ranks
    DevBuf rss, rss_partials, out_tmp, cand_terms, Ninv, tgt_terms, rss_corr, grp_buf, rep_buf, zsol, canon_tmp;
    int nstreamed = 0;    // candidates actually streamed by the last rss pass (group representatives)
    const double* rss_ptr = nullptr;  // table used by select (own buffer or the caller's)
    void release_all();
};

int rss_run(ddb_ctx* ctx, SweepState* S, double* dRss);

}  // namespace ddb
