// DataDrivenDMD Gram path (P = X X', Q = Y X', K = Q P^-1): see include/ddb.h.  Not implemented in
// this revision; the entry points report DDB_UNSUPPORTED (there is no CPU fallback).
#include "common.cuh"
using namespace ddb;
extern "C" {
int ddb_dmd_gram(ddb_ctx*, const double*, const double*, int, int64_t, double*, double*) {
    set_error("ddb_dmd_gram: not implemented yet");
    return DDB_UNSUPPORTED;
}
int ddb_dmd_operator(ddb_ctx*, const double*, const double*, int, double*) {
    set_error("ddb_dmd_operator: not implemented yet");
    return DDB_UNSUPPORTED;
}
}
