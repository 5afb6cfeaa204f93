#include "common.cuh"
namespace ddb {
void sweep_free(ddb_ctx* ctx) { (void)ctx; }
}
using namespace ddb;
#define NOT_YET(name) set_error(name ": not implemented yet"); return DDB_UNSUPPORTED
extern "C" {
int ddb_sweep(ddb_ctx*, const ddb_sweep_params*, const double*, int, int64_t) { NOT_YET("ddb_sweep"); }
int64_t ddb_rss_count(const ddb_ctx*) { return 0; }
int ddb_rss(ddb_ctx*, double*) { NOT_YET("ddb_rss"); }
int ddb_rss_buffer(ddb_ctx*, double**) { NOT_YET("ddb_rss_buffer"); }
int ddb_select(ddb_ctx*, double*, double*, int32_t*, double*, int32_t*, uint32_t*, double*, int32_t*, int32_t*) { NOT_YET("ddb_select"); }
int ddb_solve_sparse(ddb_ctx*, const double*, const double*, int, int64_t, const ddb_sweep_params*, const double*, int, double*, double*, int32_t*, double*, int32_t*, uint32_t*) { NOT_YET("ddb_solve_sparse"); }
int ddb_dmd_gram(ddb_ctx*, const double*, const double*, int, int64_t, double*, double*) { NOT_YET("ddb_dmd_gram"); }
int ddb_dmd_operator(ddb_ctx*, const double*, const double*, int, double*) { NOT_YET("ddb_dmd_operator"); }
}
