// Cell-update math shared by the recurrent kernel (K3, lstm_recurrent_tc.cu) and the fused cell epilogue of the wave
// decoder's GEMM (proj_gemm_tc.cuh): packed fp32 pairs, bare MUFU forms, shared reciprocals.
// Reference semantics: tf.keras LSTMCell / GRUCell(reset_after=True) as used by basecaller.py:24-44, 76-88.
#pragma once
#include <cstdint>

namespace rvb {
namespace cellmath {

// LSTM cell pointwise math with shared reciprocals: 5 ex2 + 3 rcp per unit instead of 5 + 5.
//   i*g      = (1/A)(1 - 2/G) = (G - 2) / (A*G)          A = 1 + e^-zi,  G = 1 + e^{2 zg}
//   f, o     : r = 1/(F*O);  f = r*O,  o = r*F            F = 1 + e^-zf,  O = 1 + e^-zo
//   o*tanh(c'): o * (Cc - 2) / Cc                         Cc = 1 + e^{2 c'}
// Arguments are clamped where the functions are saturated to < 1e-13 of their limit, which keeps every
// product below 1e27 (no fp32 overflow).
// The bare MUFU forms: every argument below is clamped first, so the range fix-ups that __expf / __fdividef wrap around
// ex2.approx / rcp.approx (an FSETP, an FSEL and two or three FMULs each -- a third of this loop's issue slots) are dead code.
__device__ __forceinline__ float ex2_raw(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_raw(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// Packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2): the cell update is issue bound, and two units share every fp32 slot.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f32x2 ex2_2(f32x2 x) { float a, b; upk(x, a, b); return pk(ex2_raw(a), ex2_raw(b)); }
__device__ __forceinline__ f32x2 rcp_2(f32x2 x) { float a, b; upk(x, a, b); return pk(rcp_raw(a), rcp_raw(b)); }
__device__ __forceinline__ f32x2 max_2(f32x2 x, float m) { float a, b; upk(x, a, b); return pk(fmaxf(a, m), fmaxf(b, m)); }
__device__ __forceinline__ f32x2 min_2(f32x2 x, float m) { float a, b; upk(x, a, b); return pk(fminf(a, m), fminf(b, m)); }

// Two units at once.  One-sided clamps are enough: e^-z -> 0 and e^2z -> 0 on the far side are harmless; the clamped
// side (sigmoid arguments >= -20, tanh arguments <= 10: both functions are within 5e-9 of their limits there, below
// fp32 resolution of the gate values) keeps A, F, O, G <= 1 + e^20, so even A*G*F*O <= 5.5e34 is finite and ONE
// reciprocal serves both quotients: 1/(A*G) = R*(F*O), 1/(F*O) = R*(A*G).  7 MUFU per unit (5 ex2 + 2 rcp).
__device__ __forceinline__ void lstm_pointwise2(f32x2 zi, f32x2 zf, f32x2 zg, f32x2 zo, f32x2 c, f32x2 &cn, f32x2 &hn) {
    constexpr float L2E = 1.4426950408889634f;
    const f32x2 one = pk(1.0f, 1.0f), mtwo = pk(-2.0f, -2.0f), nl = pk(-L2E, -L2E), l2 = pk(2.0f * L2E, 2.0f * L2E);
    const f32x2 A = add2(ex2_2(mul2(max_2(zi, -20.0f), nl)), one);
    const f32x2 F = add2(ex2_2(mul2(max_2(zf, -20.0f), nl)), one);
    const f32x2 O = add2(ex2_2(mul2(max_2(zo, -20.0f), nl)), one);
    const f32x2 G = add2(ex2_2(mul2(min_2(zg, 10.0f), l2)), one);
    const f32x2 AG = mul2(A, G), FO = mul2(F, O);
    const f32x2 R = rcp_2(mul2(AG, FO));
    const f32x2 ig = mul2(add2(G, mtwo), mul2(R, FO));        // sigmoid(zi) * tanh(zg) = (G - 2) / (A G)
    const f32x2 r = mul2(R, AG);                              // 1 / (F O)
    cn = fma2(c, mul2(r, O), ig);                             // sigmoid(zf) = r O
    const f32x2 Cc = add2(ex2_2(mul2(min_2(cn, 10.0f), l2)), one);
    hn = mul2(mul2(mul2(r, F), add2(Cc, mtwo)), rcp_2(Cc));   // sigmoid(zo) tanh(c') = r F (Cc - 2) / Cc
}
// Keras GRUCell (reset_after = True), two units at once, on the same four pre-activation columns per unit:
//   s0 = z gate, s1 = r gate, s2 = input part of the candidate, s3 = recurrent part of the candidate (incl. its bias)
//   z = sigmoid(s0), r = sigmoid(s1), hh = tanh(s2 + r * s3), h' = z * h + (1 - z) * hh
// One reciprocal serves both sigmoids (Z R <= (1 + e^20)^2, finite); 3 ex2 + 2 rcp per unit.
__device__ __forceinline__ void gru_pointwise2(f32x2 s0, f32x2 s1, f32x2 s2, f32x2 s3, f32x2 h, f32x2 &hn) {
    constexpr float L2E = 1.4426950408889634f;
    const f32x2 one = pk(1.0f, 1.0f), mtwo = pk(-2.0f, -2.0f), nl = pk(-L2E, -L2E), l2 = pk(2.0f * L2E, 2.0f * L2E);
    const f32x2 Z = add2(ex2_2(mul2(max_2(s0, -20.0f), nl)), one);
    const f32x2 R = add2(ex2_2(mul2(max_2(s1, -20.0f), nl)), one);
    const f32x2 inv = rcp_2(mul2(Z, R));
    const f32x2 zg = mul2(inv, R), rg = mul2(inv, Z);             // sigmoid(s0) = 1/Z, sigmoid(s1) = 1/R
    const f32x2 pre = fma2(rg, s3, s2);
    const f32x2 T2 = add2(ex2_2(mul2(min_2(pre, 10.0f), l2)), one);
    const f32x2 hh = mul2(add2(T2, mtwo), rcp_2(T2));             // tanh(pre) = (T2 - 2) / T2
    hn = fma2(zg, add2(h, mul2(hh, pk(-1.0f, -1.0f))), hh);       // hh + z (h - hh)
}
}  // namespace cellmath
}  // namespace rvb
