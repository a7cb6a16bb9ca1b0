// Division of a double by a small positive integer without the IEEE division sequence (rs_maps.cu: the running
// standardiser of the map observation, RADTEAM_core.py:215-265, divides by the reading count twice per agent and call).
#pragma once

// x / k for a small positive integer k through its correctly rounded reciprocal (a compile-time table) and one exact
// residual correction: q0 = x * (1/k), r = x - q0 * k (one FMA, exact), q = q0 + r * (1/k).  Markstein's theorem makes q the
// correctly rounded quotient (the IEEE division's result) for every finite x when the reciprocal is correctly rounded and
// the divisor's significand is not all ones -- true of every integer below 2^53 - 1 -- as long as nothing underflows (guarded below).  The standardiser divides by the
// reading count twice per agent and call; the two IEEE divisions were a sixth of the kernel's instructions.
constexpr int kRcpN = 4096;
struct RcpTable {
    double v[kRcpN];
    constexpr RcpTable() : v() {
        for (int i = 1; i < kRcpN; i++) v[i] = 1.0 / (double)i;
    }
};
#ifdef RS_HOST_EMU
static const RcpTable g_rcp = RcpTable();
#else
__constant__ RcpTable g_rcp = RcpTable();
#endif

__device__ __forceinline__ double div_count(double x, int k) {
    // outside the table, non-finite, or so close to the subnormal range that the residual could underflow: the IEEE division
    const double ax = fabs(x);
    if (k <= 0 || k >= kRcpN || !(ax < 1e300) || (ax < 1e-290 && ax > 0.0)) return x / (double)k;
    const double d = (double)k, rd = g_rcp.v[k];
    const double q0 = __dmul_rn(x, rd);
    const double r = __fma_rn(-q0, d, x);
    return __fma_rn(r, rd, q0);
}

