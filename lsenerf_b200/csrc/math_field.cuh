// math_field.cuh -- per-sample math of LSEField: position, L-inf scene contraction, hash-grid corner
// indices / trilinear weights, degree-4 spherical harmonics; forward and reverse mode.
//
// Reference semantics:
//   lse_nerf/lse_field.py:266-274  positions = (contract(o + d*(t0+t1)/2) + 2)/4 ; selector = all(0<p<1) ; p*=selector
//   lse_nerf/lse_field.py:58-65    scalings (fp32 floor(min_res*growth**l)) come from the module buffer
//   nerfstudio 0.3.2 HashEncoding.pytorch_fwd / hash_fn (3P, reached at lse_field.py:208,279):
//       scaled = x*scalings[l]; c = ceil, f = floor; offset = scaled - f;
//       idx = ((X*1) ^ (Y*2654435761) ^ (Z*805459861)) % T + l*T   (int64 in torch; T = 2^k and X,Y,Z >= 0, so the
//       low k bits of a wrapping uint32 product are identical -- tests/test_hash_index.py proves it)
//       corner order 0=(c,c,c) 1=(c,f,c) 2=(f,f,c) 3=(f,c,c) 4=(c,c,f) 5=(c,f,f) 6=(f,f,f) 7=(f,c,f);
//       blend weight `offset` goes to the CEIL corner.
//   nerfstudio components_from_spherical_harmonics(levels=4) on the shifted vector (d+1)/2 (lse_field.py:298-300)
#pragma once
#include "common.cuh"

namespace lse {

// world position of a sample, rounded like torch eager (separate mul / add):  o + d * ((t0+t1)/2)
LSE_HD void sample_position(const float o[3], const float d[3], float t0, float t1, float p[3])
{
    float tm = rn_div(rn_add(t0, t1), 2.0f);
    for (int i = 0; i < 3; ++i) p[i] = rn_add(o[i], rn_mul(d[i], tm));
}

// SceneContraction(order=inf) then (p+2)/4 then selector.  Returns selector; x = normalised position * selector.
// ctx: m = |p|_inf, kmax = argmax axis (first maximal axis), outside = m > 1.
struct ContractCtx { float m; int kmax; bool outside; bool sel; };

LSE_HD bool contract_normalize(const float p[3], float x[3], ContractCtx &c)
{
    float a0 = fabsf(p[0]), a1 = fabsf(p[1]), a2 = fabsf(p[2]);
    c.m = fmaxf(a0, fmaxf(a1, a2));
    c.kmax = (a0 >= a1 && a0 >= a2) ? 0 : (a1 >= a2 ? 1 : 2);
    c.outside = !(c.m <= 1.0f);
    float q[3];
    if (c.outside) {
        float s = rn_sub(2.0f, rn_div(1.0f, c.m));
        for (int i = 0; i < 3; ++i) q[i] = rn_mul(s, rn_div(p[i], c.m));
    } else {
        for (int i = 0; i < 3; ++i) q[i] = p[i];
    }
    bool sel = true;
    for (int i = 0; i < 3; ++i) {
        q[i] = rn_div(rn_add(q[i], 2.0f), 4.0f);
        sel = sel && (q[i] > 0.0f) && (q[i] < 1.0f);
    }
    c.sel = sel;
    for (int i = 0; i < 3; ++i) x[i] = sel ? q[i] : rn_mul(q[i], 0.0f);
    return sel;
}

// d(x)/d(p) applied to an upstream gradient dx -> dp
LSE_HD void contract_normalize_bwd(const float p[3], const ContractCtx &c, const float dx[3], float dp[3])
{
    if (!c.sel) { dp[0] = dp[1] = dp[2] = 0.0f; return; }
    float g[3] = {dx[0] * 0.25f, dx[1] * 0.25f, dx[2] * 0.25f};
    if (!c.outside) { dp[0] = g[0]; dp[1] = g[1]; dp[2] = g[2]; return; }
    // q = s(m) * p / m with s = 2 - 1/m  ->  q = r(m) * p, r = 2/m - 1/m^2, dr/dm = -2/m^2 + 2/m^3
    float m = c.m;
    float r = 2.0f / m - 1.0f / (m * m);
    float drdm = -2.0f / (m * m) + 2.0f / (m * m * m);
    float pg = p[0] * g[0] + p[1] * g[1] + p[2] * g[2];
    for (int i = 0; i < 3; ++i) dp[i] = r * g[i];
    float sgn = p[c.kmax] >= 0.0f ? 1.0f : -1.0f;
    dp[c.kmax] += pg * drdm * sgn;
}

// ---- hash grid -----------------------------------------------------------------------------------------------
constexpr uint32_t kPrimeY = 2654435761u;
constexpr uint32_t kPrimeZ = 805459861u;

struct HashCell {
    uint32_t idx[8];  // row inside the level's sub-table (without the l*T offset)
    float ox, oy, oz;  // blend offsets (weight of the ceil corner)
};

LSE_HD void hash_cell(const float x[3], float scale, uint32_t mask, HashCell &h)
{
    float sx = rn_mul(x[0], scale), sy = rn_mul(x[1], scale), sz = rn_mul(x[2], scale);
    float fxf = floorf(sx), fyf = floorf(sy), fzf = floorf(sz);
    uint32_t fx = (uint32_t)(int32_t)fxf, fy = (uint32_t)(int32_t)fyf, fz = (uint32_t)(int32_t)fzf;
    uint32_t cx = (uint32_t)(int32_t)ceilf(sx), cy = (uint32_t)(int32_t)ceilf(sy), cz = (uint32_t)(int32_t)ceilf(sz);
    h.ox = rn_sub(sx, fxf); h.oy = rn_sub(sy, fyf); h.oz = rn_sub(sz, fzf);
    uint32_t cyp = cy * kPrimeY, fyp = fy * kPrimeY, czp = cz * kPrimeZ, fzp = fz * kPrimeZ;
    h.idx[0] = (cx ^ cyp ^ czp) & mask;
    h.idx[1] = (cx ^ fyp ^ czp) & mask;
    h.idx[2] = (fx ^ fyp ^ czp) & mask;
    h.idx[3] = (fx ^ cyp ^ czp) & mask;
    h.idx[4] = (cx ^ cyp ^ fzp) & mask;
    h.idx[5] = (cx ^ fyp ^ fzp) & mask;
    h.idx[6] = (fx ^ fyp ^ fzp) & mask;
    h.idx[7] = (fx ^ cyp ^ fzp) & mask;
}

// the 8 trilinear weights in corner order (products evaluated in the reference's association is irrelevant for the
// weights themselves; the blend below reproduces the reference's nested lerps)
LSE_HD void hash_blend(const float F[8][2], const HashCell &h, float out[2])
{
    float ox = h.ox, oy = h.oy, oz = h.oz;
    for (int f = 0; f < 2; ++f) {
        float f03 = F[0][f] * ox + F[3][f] * (1.0f - ox);
        float f12 = F[1][f] * ox + F[2][f] * (1.0f - ox);
        float f56 = F[5][f] * ox + F[6][f] * (1.0f - ox);
        float f47 = F[4][f] * ox + F[7][f] * (1.0f - ox);
        float f0312 = f03 * oy + f12 * (1.0f - oy);
        float f4756 = f47 * oy + f56 * (1.0f - oy);
        out[f] = f0312 * oz + f4756 * (1.0f - oz);
    }
}

LSE_HD void hash_weights(const HashCell &h, float w[8])
{
    float ox = h.ox, oy = h.oy, oz = h.oz, nx = 1.0f - ox, ny = 1.0f - oy, nz = 1.0f - oz;
    w[0] = ox * oy * oz; w[3] = nx * oy * oz; w[1] = ox * ny * oz; w[2] = nx * ny * oz;
    w[4] = ox * oy * nz; w[7] = nx * oy * nz; w[5] = ox * ny * nz; w[6] = nx * ny * nz;
}

// d(out)/d(offset) for both features contracted with the upstream gradient g[2] -> d(scaled xyz)
LSE_HD void hash_blend_dx(const float F[8][2], const HashCell &h, const float g[2], float d[3])
{
    float ox = h.ox, oy = h.oy, oz = h.oz;
    d[0] = d[1] = d[2] = 0.0f;
    for (int f = 0; f < 2; ++f) {
        float f03 = F[0][f] * ox + F[3][f] * (1.0f - ox), d03 = F[0][f] - F[3][f];
        float f12 = F[1][f] * ox + F[2][f] * (1.0f - ox), d12 = F[1][f] - F[2][f];
        float f56 = F[5][f] * ox + F[6][f] * (1.0f - ox), d56 = F[5][f] - F[6][f];
        float f47 = F[4][f] * ox + F[7][f] * (1.0f - ox), d47 = F[4][f] - F[7][f];
        float f0312 = f03 * oy + f12 * (1.0f - oy);
        float f4756 = f47 * oy + f56 * (1.0f - oy);
        float dxv = (d03 * oy + d12 * (1.0f - oy)) * oz + (d47 * oy + d56 * (1.0f - oy)) * (1.0f - oz);
        float dyv = (f03 - f12) * oz + (f47 - f56) * (1.0f - oz);
        float dzv = f0312 - f4756;
        d[0] += g[f] * dxv; d[1] += g[f] * dyv; d[2] += g[f] * dzv;
    }
}

// ---- spherical harmonics, 16 components --------------------------------------------------------------------
LSE_HD void sh16(const float v[3], float o[16])
{
    float x = v[0], y = v[1], z = v[2], xx = x * x, yy = y * y, zz = z * z;
    o[0] = 0.28209479177387814f;
    o[1] = 0.4886025119029199f * y;
    o[2] = 0.4886025119029199f * z;
    o[3] = 0.4886025119029199f * x;
    o[4] = 1.0925484305920792f * x * y;
    o[5] = 1.0925484305920792f * y * z;
    o[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
    o[7] = 1.0925484305920792f * x * z;
    o[8] = 0.5462742152960396f * (xx - yy);
    o[9] = 0.5900435899266435f * y * (3.0f * xx - yy);
    o[10] = 2.890611442640554f * x * y * z;
    o[11] = 0.4570457994644658f * y * (5.0f * zz - 1.0f);
    o[12] = 0.3731763325901154f * z * (5.0f * zz - 3.0f);
    o[13] = 0.4570457994644658f * x * (5.0f * zz - 1.0f);
    o[14] = 1.445305721320277f * z * (xx - yy);
    o[15] = 0.5900435899266435f * x * (xx - 3.0f * yy);
}

LSE_HD void sh16_bwd(const float v[3], const float g[16], float dv[3])
{
    float x = v[0], y = v[1], z = v[2], xx = x * x, yy = y * y, zz = z * z;
    const float c1 = 0.4886025119029199f, c2 = 1.0925484305920792f, c6 = 0.9461746957575601f, c8 = 0.5462742152960396f,
                c9 = 0.5900435899266435f, c10 = 2.890611442640554f, c11 = 0.4570457994644658f,
                c12 = 0.3731763325901154f, c14 = 1.445305721320277f;
    float dx = 0.0f, dy = 0.0f, dz = 0.0f;
    dy += g[1] * c1; dz += g[2] * c1; dx += g[3] * c1;
    dx += g[4] * c2 * y; dy += g[4] * c2 * x;
    dy += g[5] * c2 * z; dz += g[5] * c2 * y;
    dz += g[6] * c6 * 2.0f * z;
    dx += g[7] * c2 * z; dz += g[7] * c2 * x;
    dx += g[8] * c8 * 2.0f * x; dy -= g[8] * c8 * 2.0f * y;
    dx += g[9] * c9 * y * 6.0f * x; dy += g[9] * c9 * (3.0f * xx - 3.0f * yy);
    dx += g[10] * c10 * y * z; dy += g[10] * c10 * x * z; dz += g[10] * c10 * x * y;
    dy += g[11] * c11 * (5.0f * zz - 1.0f); dz += g[11] * c11 * y * 10.0f * z;
    dz += g[12] * c12 * (15.0f * zz - 3.0f);
    dx += g[13] * c11 * (5.0f * zz - 1.0f); dz += g[13] * c11 * x * 10.0f * z;
    dx += g[14] * c14 * z * 2.0f * x; dy -= g[14] * c14 * z * 2.0f * y; dz += g[14] * c14 * (xx - yy);
    dx += g[15] * c9 * (3.0f * xx - 3.0f * yy); dy -= g[15] * c9 * x * 6.0f * y;
    dv[0] = dx; dv[1] = dy; dv[2] = dz;
}

LSE_HD float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace lse
