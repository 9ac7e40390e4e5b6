// Integer-pipe peak probes: the measured denominator of the DP roofline.
// Each thread runs kChains independent dependency chains so the issue rate, not the latency,
// is what is measured; one "lane-op" = one thread executing one instruction.
#pragma once
#include "common.cuh"

namespace ovl {

constexpr int kProbeChains = 8;
constexpr int kProbeUnroll = 16;
constexpr int kProbeThreads = 256;

template <int KIND>
__global__ void __launch_bounds__(kProbeThreads) int_probe_kernel(uint32_t* out, int iters, uint32_t a0, uint32_t b0, uint32_t one) {
    uint32_t v[kProbeChains], w[kProbeChains], x[kProbeChains], y[kProbeChains], z[kProbeChains];
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) {
        v[c] = a0 + threadIdx.x + c; w[c] = b0 ^ (c * 0x01010101u);
        x[c] = a0 * (c + 3); y[c] = b0 + 7 * c + threadIdx.x; z[c] = (a0 ^ b0) + c;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kProbeUnroll; ++u) {
#pragma unroll
            for (int c = 0; c < kProbeChains; ++c) {
                if (KIND == 0) {            // IADD3
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(v[c]) : "r"(w[c]));
                } else if (KIND == 1) {     // IMAD
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(one), "r"(w[c]));
                } else if (KIND == 2) {     // VIMNMX.S32
                    asm volatile("max.s32 %0, %0, %1;" : "+r"(v[c]) : "r"(w[c]));
                    w[c] += 0;              // keep operands live
                } else if (KIND == 3) {     // VIADDMNMX.S16x2
                    v[c] = __viaddmax_s16x2(v[c], b0, w[c]);
                } else if (KIND == 4) {     // DP inner-loop mix: PRMT + IMAD + 2x VIADDMNMX.S16x2
                    // every op depends on the chain value so ptxas cannot hoist any of them
                    uint32_t sc;
                    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(sc) : "r"(a0), "r"(b0), "r"(v[c]));
                    uint32_t d;
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(sc), "r"(one), "r"(w[c]));
                    uint32_t t1 = __viaddmax_s16x2(w[c], b0, d);
                    v[c] = __viaddmax_s16x2(v[c], a0, t1);
                } else if (KIND == 7) {     // LOP3 and IMAD on independent chains: do ALU and FMA pipes co-issue?
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(a0), "r"(b0));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[c]) : "r"(one), "r"(b0));
                } else if (KIND == 8) {     // VIMNMX3 + IMAD with all-distinct register operands: register-file pressure
                    v[c] = __vimin3_u16x2(v[c], x[c], y[c]);
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[c]) : "r"(one), "r"(z[c]));
                } else if (KIND == 9) {     // form-1 column with distinct registers per chain (PRMT, IMAD, 2x VIADDMNMX)
                    uint32_t dc;
                    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(dc) : "r"(x[c]), "r"(y[c]), "r"(v[c]));
                    uint32_t a1;
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a1) : "r"(dc), "r"(one), "r"(z[c]));
                    uint32_t t1 = __viaddmin_u16x2(w[c], b0, a1);
                    v[c] = __viaddmin_u16x2(v[c], a0, t1);
                } else if (KIND == 10) {    // form-2 DP column per chain (PRMT, 3x IMAD, VIMNMX3)
                    uint32_t dc;
                    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(dc) : "r"(x[c]), "r"(y[c]), "r"(v[c]));
                    uint32_t a1, a2, a3;
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a1) : "r"(dc), "r"(one), "r"(z[c]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a2) : "r"(w[c]), "r"(one), "r"(b0));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a3) : "r"(v[c]), "r"(one), "r"(a0));
                    v[c] = __vimin3_u16x2(a1, a2, a3);
                } else if (KIND == 11) {    // 2 ALU (PRMT, VIMNMX3) + 2 IMAD per chain: balanced pipes
                    uint32_t dc;
                    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(dc) : "r"(x[c]), "r"(y[c]), "r"(v[c]));
                    uint32_t a1, a3;
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a1) : "r"(dc), "r"(one), "r"(z[c]));
                    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a3) : "r"(v[c]), "r"(one), "r"(a0));
                    v[c] = __vimin3_u16x2(a1, w[c], a3);
                } else if (KIND == 12) {    // PRMT + VIADDMNMX on independent chains (ALU only)
                    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(x[c]), "r"(y[c]));
                    w[c] = __viaddmin_u16x2(w[c], b0, z[c]);
                } else if (KIND == 13) {    // VIADDMNMX + IMAD on independent chains
                    v[c] = __viaddmin_u16x2(v[c], b0, x[c]);
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[c]) : "r"(one), "r"(z[c]));
                } else if (KIND == 14) {    // PRMT + IMAD on independent chains
                    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(x[c]), "r"(y[c]));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[c]) : "r"(one), "r"(z[c]));
                } else if (KIND == 15) {    // 2x VIADDMNMX + IMAD (no PRMT), independent chains
                    v[c] = __viaddmin_u16x2(v[c], b0, x[c]);
                    y[c] = __viaddmin_u16x2(y[c], a0, x[c]);
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[c]) : "r"(one), "r"(z[c]));
                } else if (KIND == 16) {    // form-1 column but the diagonal add on the ALU (IADD3) instead of IMAD
                    uint32_t dc;
                    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(dc) : "r"(x[c]), "r"(y[c]), "r"(v[c]));
                    uint32_t a1 = dc + z[c];
                    uint32_t t1 = __viaddmin_u16x2(w[c], b0, a1);
                    v[c] = __viaddmin_u16x2(v[c], a0, t1);
                } else if (KIND == 17) {    // 2-input packed min (__vminu2): SASS is VIMNMX3.U16x2 with a repeated operand
                    v[c] = __vminu2(v[c], w[c]);
                    w[c] += 0;
                } else if (KIND == 18) {    // VIMNMX3.U16x2 alone (3-input packed min)
                    v[c] = __vimin3_u16x2(v[c], w[c], x[c]);
                } else if (KIND == 20) {    // VIADDMNMX.U16x2 alone (the kernel's fused add+min)
                    v[c] = __viaddmin_u16x2(v[c], b0, w[c]);
                } else if (KIND == 21) {    // VIADDMNMX.U16x2 with an IMMEDIATE addend: two register sources instead of three
                    v[c] = __viaddmin_u16x2(v[c], 0x70007000u, w[c]);
                } else if (KIND == 22) {    // form-1 column with immediate gap costs: PRMT, IMAD.IADD, 2x VIADDMNMX(imm)
                    uint32_t dc;
                    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(dc) : "r"(x[c]), "r"(y[c]), "r"(v[c]));
                    uint32_t a1 = dc + z[c];
                    uint32_t t1 = __viaddmin_u16x2(w[c], 0x70007000u, a1);
                    v[c] = __viaddmin_u16x2(v[c], 0x70007000u, t1);
                } else if (KIND == 23) {    // IMAD with an immediate multiplier (two register sources)
                    asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(v[c]) : "r"(w[c]));
                } else if (KIND == 24) {    // PRMT with a repeated source (two distinct registers)
                    asm volatile("prmt.b32 %0, %0, %0, %1;" : "+r"(v[c]) : "r"(w[c]));
                } else if (KIND == 25) {    // LOP3 with an immediate (two register sources)
                    asm volatile("lop3.b32 %0, %0, %1, 0x0f0f0f0f, 0x96;" : "+r"(v[c]) : "r"(w[c]));
                } else if (KIND == 26) {    // fp16x2 min alone (HMNMX2): which pipe, what rate?
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(w[c]));
                } else if (KIND == 27) {    // HMNMX2 + LOP3 on independent chains
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(x[c]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[c]) : "r"(a0), "r"(b0));
                } else if (KIND == 28) {    // HMNMX2 + IMAD on independent chains
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(x[c]));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[c]) : "r"(one), "r"(b0));
                } else if (KIND == 29) {    // HMNMX2 + VIADDMNMX.U16x2 on independent chains
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(x[c]));
                    w[c] = __viaddmin_u16x2(w[c], 0x3e003e00u, z[c]);
                } else if (KIND == 30) {    // HMNMX2 + PRMT on independent chains
                    asm volatile("min.f16x2 %0, %0, %1;" : "+r"(v[c]) : "r"(x[c]));
                    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(w[c]) : "r"(y[c]), "r"(z[c]));
                } else if (KIND == 5) {     // PRMT
                    asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[c]) : "r"(w[c]), "r"(b0));
                } else {                    // LOP3
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[c]) : "r"(w[c]), "r"(b0));
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < kProbeChains; ++c) acc ^= v[c] + w[c] + (KIND >= 8 ? x[c] ^ y[c] ^ z[c] : 0u);
    if (acc == 0x12345678u) out[0] = acc;    // practically never; keeps the chains alive
}

inline int probe_ops_per_iter(int kind) {
    if (kind == 10) return 5;
    if (kind == 4 || kind == 9 || kind == 11 || kind == 16 || kind == 22) return 4;
    if (kind == 15) return 3;
    return (kind == 7 || kind == 8 || kind == 12 || kind == 13 || kind == 14 || (kind >= 27 && kind <= 30)) ? 2 : 1;
}

}  // namespace ovl
