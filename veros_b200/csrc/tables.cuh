// Metric tables shared by the kernels.  setup_kernel builds them once per op call in the workspace
// (a few KB to a few MB, L1/L2 resident); no kernel on the hot path divides by a grid metric.
// Every Divisor holds the divisor and its correctly rounded reciprocal (strict.cuh), so the strict
// kernels get exact quotients in three FP64 instructions and the fast kernel a plain reciprocal.
#pragma once

#include "common.cuh"
#include "strict.cuh"

namespace vb {

struct LevTab {           // per level k
    strict::Divisor d4zt;  // 4 dzt[k]
    strict::Divisor dzt;   // dzt[k]
    strict::Divisor dzw;   // dzw[k]
    double dt_dzw;         // dt_tracer / dzw[k]   (IEEE division, diffusion.py:155)
    double pabs;           // |zt[k]|
};
struct RowTab {            // per row j
    strict::Divisor dyu, cost, d4ytc, cdyt;  // dyu[j], cost[j], (4 dyt[j]) cost[j], cost[j] dyt[j]
    double cosu, facty;                       // cosu[j], cosu[j] dyu[j]
};
struct XTab {              // per plane i
    strict::Divisor d4xt;  // 4 dxt[i]
    double dxu, pad;
};
struct CellTab {           // per (i, j)
    strict::Divisor cdxu;  // cost[j] dxu[i]
    strict::Divisor cdxt;  // cost[j] dxt[i]
};

struct Tables {
    const LevTab* lev;
    const RowTab* row;
    const XTab* xt;
    const CellTab* cell;
    unsigned int* counters;  // 4 work counters, zeroed by setup_kernel (dynamic chunk scheduling of the slope kernel)
};

__host__ __device__ inline size_t tables_doubles(int N, int M, int nz) {
    return (size_t)nz * (sizeof(LevTab) / 8) + (size_t)M * (sizeof(RowTab) / 8) + (size_t)N * (sizeof(XTab) / 8) +
           (size_t)N * M * (sizeof(CellTab) / 8) + 2;
}

__host__ __device__ inline Tables tables_at(double* base, int N, int M, int nz) {
    Tables t;
    LevTab* lev = reinterpret_cast<LevTab*>(base);
    RowTab* row = reinterpret_cast<RowTab*>(lev + nz);
    XTab* xt = reinterpret_cast<XTab*>(row + M);
    CellTab* cell = reinterpret_cast<CellTab*>(xt + N);
    t.lev = lev;
    t.row = row;
    t.xt = xt;
    t.cell = cell;
    t.counters = reinterpret_cast<unsigned int*>(cell + (size_t)N * M);
    return t;
}

__device__ __forceinline__ strict::Divisor ld_div(const strict::Divisor* p) {
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return strict::Divisor{v.x, v.y};
}

// defined in iso_pre.cu
// `zero` / `nzero`: unsigned ints to clear in the same launch (the fused kernel's queue and counters)
void launch_setup_tables(cudaStream_t s, const Grid& g, double dt_tracer, double* tables, unsigned int* zero = nullptr,
                         int nzero = 0);

}  // namespace vb
