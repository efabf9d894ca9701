// ppf_math.cuh -- device arithmetic of the PPF hot path, written for sm_100a.
//
// Bit-exactness contract.  The parity target is the reference's CUDA path as
// nvcc 12.9 compiles it with --ftz=true --prec-div=false --prec-sqrt=false
// (pcl/alignment/CMakeLists.txt:71).  Both nvcc and ptxas are free to fuse an
// unqualified mul+add, so every floating-point operation below that must match
// the reference is written with an explicit, non-contractable intrinsic
// (__fmul_rn / __fadd_rn / __fmaf_rn) or inline PTX (sqrt.approx.ftz,
// div.full.ftz), in the order the reference's SASS performs them.  This file
// must be compiled with the same three flags so that libdevice's acosf /
// atan2f / sinf / cosf expand to the same instruction sequences.
//
// Reference semantics restated here (file:line in /root/reference/pcl/alignment):
//   dot/norm            src/cuda/kernel.cu:51-65
//   compute_ppf         src/cuda/kernel.cu:109-122
//   quant_downf         src/cuda/kernel.cu:90-92   (closed form, see quant_bin)
//   hash (FNV-1a)       src/cuda/kernel.cu:23-30
//   trans/rot*/mat4f_*  src/cuda/kernel.cu:170-252
//   trans_model_scene   src/cuda/kernel.cu:302-349
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace ppf {

constexpr int   kNAngle     = 30;                       // kernel.h:15
constexpr int   kNAlphaBins = 31;                       // alpha_idx in [0,30] (kernel.cu:341-342)
constexpr int   kAngleCells = 17;                       // k in 0..15, 16 = NaN feature
constexpr int   kCellsPerDist = kAngleCells * kAngleCells * kAngleCells;
constexpr uint32_t kNoBucket = 0xFFFFFFFFu;

// D_ANGLE0 = (2.0f*float(CUDART_PI_F))/float(N_ANGLE)   (kernel.h:16) == 0x3E567750
__host__ __device__ __forceinline__ constexpr float d_angle0() {
    return (2.0f * float(CUDART_PI_F)) / float(kNAngle);
}

// ---- primitive ops -----------------------------------------------------------
__device__ __forceinline__ float sqrt_approx_ftz(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float div_full_ftz(float a, float b) {
    float r;
    asm("div.full.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// kernel.cu:51-53 as compiled: t = y*y'; t = fma(x,x',t); t = fma(z,z',t)
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
    return __fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by)));
}
// kernel.cu:55-57 as compiled: t = y*y'; fma x; fma z; fma w
__device__ __forceinline__ float dot4(float ax, float ay, float az, float aw,
                                      float bx, float by, float bz, float bw) {
    return __fmaf_rn(aw, bw, __fmaf_rn(az, bz, __fmaf_rn(ax, bx, __fmul_rn(ay, by))));
}

// ---- quantiser ---------------------------------------------------------------
// The reference quantises with x - fmodf(x, step) (kernel.cu:90-92).  fmodf is
// exact, so the result is RN(k*step) with k = floor(x/step) evaluated exactly.
// quant_bin returns that k without the 80-instruction fmodf: an approximate
// quotient is corrected with one exact fma remainder.  Valid for 0 <= x, k < 2^22.
__device__ __forceinline__ int quant_bin(float x, float step, float inv_step) {
    int k = __float2int_rz(__fmul_rn(x, inv_step));
    float r = __fmaf_rn(-(float)k, step, x);          // exact sign / exact compare against step
    if (r < 0.0f) k -= 1;
    else if (r >= step) k += 1;
    return k;
}
// The float the reference stores for bin k: RN(k*step).
__device__ __forceinline__ float quant_value(int k, float step) { return __fmul_rn((float)k, step); }

// ---- FNV-1a over raw bytes, bytes sign-extended (char is signed: PTX ld.s8) ----
__host__ __device__ __forceinline__ uint32_t fnv1a_word(uint32_t h, uint32_t w) {
#pragma unroll
    for (int b = 0; b < 4; b++) {
        int32_t c = (int32_t)(int8_t)(w >> (8 * b));
        h ^= (uint32_t)c;
        h *= 16777619u;
    }
    return h;
}
__host__ __device__ __forceinline__ uint32_t fnv1a_4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = 2166136261u;                         // kernel.h:22
    h = fnv1a_word(h, a); h = fnv1a_word(h, b); h = fnv1a_word(h, c); h = fnv1a_word(h, d);
    return h;
}
__host__ __device__ __forceinline__ uint32_t fnv1a_3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = 2166136261u;
    h = fnv1a_word(h, a); h = fnv1a_word(h, b); h = fnv1a_word(h, c);
    return h;
}

// ---- pair feature --------------------------------------------------------------
struct PointN {            // one cloud point as the kernels hold it
    float x, y, z;         // position
    float nx, ny, nz;      // normal (not assumed unit: VoxelGrid averages them)
    float nn;              // norm(n) = sqrt.approx(dot(n,n))
};

__device__ __forceinline__ float norm3(float x, float y, float z) {
    return sqrt_approx_ftz(dot3(x, y, z, x, y, z));
}

struct FeatureBins { int kd, k1, k2, k3; float f1; };

// compute_ppf (kernel.cu:109-122) + disc_feature (kernel.cu:94-100), returning
// bin indices instead of the quantised floats.  Angle bin 16 == NaN feature.
__device__ __forceinline__ int angle_bin(float c) {
    float a = acosf(c);
    if (a != a) return 16;
    const float D = d_angle0();
    return quant_bin(a, D, 4.7746482f);
}

__device__ __forceinline__ FeatureBins pair_feature_bins(const PointN &r, const PointN &o,
                                                         float d_dist, float inv_d_dist) {
    FeatureBins fb;
    float dx = __fsub_rn(o.x, r.x), dy = __fsub_rn(o.y, r.y), dz = __fsub_rn(o.z, r.z);
    float nd = norm3(dx, dy, dz);
    fb.f1 = nd;
    float c1 = div_full_ftz(dot3(r.nx, r.ny, r.nz, dx, dy, dz), __fmul_rn(nd, r.nn));
    float c2 = div_full_ftz(dot3(o.nx, o.ny, o.nz, dx, dy, dz), __fmul_rn(nd, o.nn));
    float c3 = div_full_ftz(dot3(r.nx, r.ny, r.nz, o.nx, o.ny, o.nz), __fmul_rn(r.nn, o.nn));
    fb.k1 = angle_bin(c1);
    fb.k2 = angle_bin(c2);
    fb.k3 = angle_bin(c3);
    // distance bin; NaN / Inf / huge distances report a bin outside every table
    if (!(nd < 3.0e38f)) fb.kd = -1;                           // isnan(.x) after quantising -> key 0
    else {
        float q = __fmul_rn(nd, inv_d_dist);
        fb.kd = (q < 4.0e6f) ? quant_bin(nd, d_dist, inv_d_dist) : 0x7FFFFFFF;
    }
    return fb;
}

__device__ __forceinline__ uint32_t angle_bits(int k) {
    return (k >= 16) ? 0x7FFFFFFFu : __float_as_uint(quant_value(k, d_angle0()));
}
// The reference's hash key for a quantised feature (ppf_hash_kernel, kernel.cu:460-477).
__device__ __forceinline__ uint32_t feature_key(int kd, int k1, int k2, int k3, float d_dist) {
    if (kd < 0) return 0u;                                      // isnan(ppf.x) -> 0
    uint32_t b0 = __float_as_uint(quant_value(kd, d_dist));
    return fnv1a_4(b0, angle_bits(k1), angle_bits(k2), angle_bits(k3));
}
__device__ __forceinline__ uint32_t cell_index(int kd, int k1, int k2, int k3) {
    return (uint32_t)(((kd * kAngleCells + k1) * kAngleCells + k2) * kAngleCells + k3);
}

// ---- local frames (trans_model_scene, kernel.cu:302-336) -----------------------
// 4x4 helpers that keep every multiplication by a structural zero, exactly as
// the compiled reference does (its SASS carries FFMA RZ,... terms), so that
// signed zeros come out identical.
struct Mat4 { float m[4][4]; };

__device__ __forceinline__ void mat4_zero(Mat4 &T) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) T.m[i][j] = 0.0f;
}
// mat4f_mul (kernel.cu:211-223): C[i][j] = fma chain over k starting from +0
__device__ __forceinline__ void mat4_mul(const Mat4 &A, const Mat4 &B, Mat4 &C) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float acc = 0.0f;
#pragma unroll
            for (int k = 0; k < 4; k++) acc = __fmaf_rn(A.m[i][k], B.m[k][j], acc);
            C.m[i][j] = acc;
        }
}
__device__ __forceinline__ void mat4_trans(float x, float y, float z, Mat4 &T) {   // kernel.cu:170-179
    mat4_zero(T);
    T.m[0][0] = 1; T.m[1][1] = 1; T.m[2][2] = 1; T.m[3][3] = 1;
    T.m[0][3] = x; T.m[1][3] = y; T.m[2][3] = z;
}
__device__ __forceinline__ void mat4_rotx(float th, Mat4 &T) {                     // kernel.cu:181-189
    mat4_zero(T);
    T.m[0][0] = 1;
    T.m[1][1] = cosf(th);
    T.m[2][1] = sinf(th);
    T.m[1][2] = __fmul_rn(-1.0f, T.m[2][1]);
    T.m[2][2] = T.m[1][1];
    T.m[3][3] = 1;
}
__device__ __forceinline__ void mat4_roty(float th, Mat4 &T) {                     // kernel.cu:191-199
    mat4_zero(T);
    T.m[0][0] = cosf(th);
    T.m[0][2] = sinf(th);
    T.m[1][1] = 1;
    T.m[2][0] = __fmul_rn(-1.0f, T.m[0][2]);
    T.m[2][2] = T.m[0][0];
    T.m[3][3] = 1;
}
__device__ __forceinline__ void mat4_rotz(float th, Mat4 &T) {                     // kernel.cu:201-209
    mat4_zero(T);
    T.m[0][0] = cosf(th);
    T.m[1][0] = sinf(th);
    T.m[0][1] = __fmul_rn(-1.0f, T.m[1][0]);
    T.m[1][1] = T.m[0][0];
    T.m[2][2] = 1;
    T.m[3][3] = 1;
}
// mat4f_vmul row (kernel.cu:234-242) with b = (x, y, z, 1)
__device__ __forceinline__ float mat4_row_apply(const float row[4], float x, float y, float z) {
    return dot4(row[0], row[1], row[2], row[3], x, y, z, 1.0f);
}

// Rotation angles of the local frame of a reference point (kernel.cu:313-316;
// same values as compute_rot_angles, kernel.cu:352-369).
__device__ __forceinline__ void frame_angles(float nx, float ny, float nz, float &roty, float &rotz) {
    Mat4 Ry;
    roty = atan2f(nz, nx);
    mat4_roty(roty, Ry);
    float tx = mat4_row_apply(Ry.m[0], nx, ny, nz);
    float ty = mat4_row_apply(Ry.m[1], nx, ny, nz);
    rotz = __fmul_rn(-1.0f, atan2f(ty, tx));
}
// T_g = Rz * Ry * Trans(-p)   (kernel.cu:310-318 / 382-388)
__device__ __forceinline__ void frame_from_angles(float px, float py, float pz, float roty, float rotz,
                                                  Mat4 &Tg) {
    Mat4 Tr, Ry, Rz, Tmp;
    mat4_trans(__fmul_rn(-1.0f, px), __fmul_rn(-1.0f, py), __fmul_rn(-1.0f, pz), Tr);
    mat4_roty(roty, Ry);
    mat4_rotz(rotz, Rz);
    mat4_mul(Rz, Ry, Tmp);
    mat4_mul(Tmp, Tr, Tg);
}

// The two rows of T_g that voting needs (y and z of T_g * (q,1)).
struct FrameYZ { float y[4]; float z[4]; };

__device__ __forceinline__ FrameYZ frame_yz(float px, float py, float pz, float nx, float ny, float nz) {
    float ry, rz;
    frame_angles(nx, ny, nz, ry, rz);
    Mat4 Tg;
    frame_from_angles(px, py, pz, ry, rz, Tg);
    FrameYZ f;
#pragma unroll
    for (int j = 0; j < 4; j++) { f.y[j] = Tg.m[1][j]; f.z[j] = Tg.m[2][j]; }
    return f;
}
// (T_g * (q,1)).yz   -- kernel.cu:330-336
__device__ __forceinline__ void frame_apply_yz(const FrameYZ &f, float qx, float qy, float qz,
                                               float &uy, float &uz) {
    uy = mat4_row_apply(f.y, qx, qy, qz);
    uz = mat4_row_apply(f.z, qx, qy, qz);
}

// ---- alpha ------------------------------------------------------------------------
constexpr int      kThetaBits = 20;
constexpr uint32_t kThetaMask = (1u << kThetaBits) - 1u;
constexpr uint32_t kThetaHalf = 1u << (kThetaBits - 1);
constexpr uint32_t kAlphaGuard = 56;                  // in units of theta LSB * 30 (error budget: DESIGN.md)
constexpr float    kTinyUV    = 1.0e-15f;

// 20-bit binary angle of (y,z) in the plane orthogonal to the normal axis.
// bit 31 of the result flags vectors whose direction is numerically meaningless
// (|u| ~ 0 or non-finite): those pairs always take the exact path.
__device__ __forceinline__ uint32_t theta_code(float y, float z) {
    float m = fmaxf(fabsf(y), fabsf(z));
    uint32_t slow = (!(m >= kTinyUV) || !(m < 3.0e38f)) ? 1u : 0u;
    float th = atan2f(z, y);
    int a = __float2int_rn(th * (float)((double)(1u << kThetaBits) / 6.283185307179586));
    return (((uint32_t)a) & kThetaMask) | (slow << 31);
}

// Exact alpha bin, as trans_model_scene computes it (kernel.cu:338-342) given
// u = (T_mg m_i).yz and v = (T_sg s_i).yz:  cross.x = uy*vz - uz*vy compiled as
// fma(uy, vz, -(uz*vy));  dot = fma(uz, vz, fma(uy, vy, +0)).
__device__ __forceinline__ uint32_t alpha_bin_exact(float uy, float uz, float vy, float vz) {
    float cr = __fmaf_rn(uy, vz, -__fmul_rn(uz, vy));
    float dt = __fmaf_rn(uz, vz, __fmaf_rn(uy, vy, 0.0f));
    float alpha = atan2f(cr, dt);
    float a2 = __fadd_rn(alpha, CUDART_PI_F);
    if (a2 != a2) return 0u;                          // lrintf(NaN) -> 0
    return (uint32_t)quant_bin(a2, d_angle0(), 4.7746482f);
}

// ---- packed voting payload ------------------------------------------------------------------
// bucket entry (model pair):  [theta_u : 20 | slow : 1 | m_r - chunk_base : 11]
// hit word     (scene pair):  [(theta_v + half) mod 2^20 : 20 | slow : 1 | 0 : 11]
// theta sits in the TOP bits so that a plain 32-bit subtraction wraps modulo one full turn.
constexpr int      kLocBits   = 11;
constexpr uint32_t kLocMask   = (1u << kLocBits) - 1u;
constexpr uint32_t kSlowBit   = 1u << kLocBits;
constexpr int      kThetaShift = 32 - kThetaBits;                 // 12
constexpr uint32_t kLowOnes   = (1u << kThetaShift) - 1u;         // 0xFFF
// The subtraction below leaves junk < 2^12 under the angle, which biases the measured in-bin
// fraction by [0, 30) guard units; the lower guard edge is widened by 30 to stay conservative.
constexpr uint32_t kGuardLoA  = (kAlphaGuard + kNAngle) << kThetaShift;
constexpr uint32_t kGuardLoB  = kAlphaGuard << kThetaShift;

__host__ __device__ __forceinline__ uint32_t pack_entry(uint32_t loc, uint32_t theta_code_u) {
    return ((theta_code_u & kThetaMask) << kThetaShift) | ((theta_code_u >> 31) << kLocBits) | loc;
}
__host__ __device__ __forceinline__ uint32_t pack_hit_theta(uint32_t theta_code_v) {
    return ((((theta_code_v & kThetaMask) + kThetaHalf) & kThetaMask) << kThetaShift) | ((theta_code_v >> 31) << kLocBits);
}
// Fast alpha bin.  t = (theta_v - theta_u + half) mod 2^20 = (alpha + pi)/2pi * 2^20 and
// bin = floor(30 t / 2^20).  `hit_ones` is the hit word with its low 12 bits set, so subtracting the
// whole entry never borrows out of the low field and the top 20 bits of the difference are exactly t;
// 30 * difference then has the bin in its high word and the in-bin fraction (<< 12) in its low word.
// Returns true when the bin is provably the reference's: fraction outside the guard band around both
// bin edges and the entry not flagged slow.
__device__ __forceinline__ bool alpha_bin_fast(uint32_t hit_ones, uint32_t entry, uint32_t &bin) {
    const uint32_t d = hit_ones - entry;
    const unsigned long long p = (unsigned long long)d * (unsigned long long)kNAngle;   // one IMAD.WIDE
    bin = (uint32_t)(p >> 32);                            // always in [0, 29]: safe to index with
    const uint32_t lo = (uint32_t)p - kGuardLoA;
    return (lo < (0u - kGuardLoA - kGuardLoB)) && !(entry & kSlowBit);
}
// Same arithmetic for the optimistic hot loop: returns the "margin" (in-bin fraction << 12, shifted
// by the lower guard edge, modulo 2^32); the fast bin is provably right iff margin < kGuardSpan, so
// the margins of a batch can be max-reduced and tested once.  bin is a valid index in every case.
constexpr uint32_t kGuardSpan = 0u - kGuardLoA - kGuardLoB;
__device__ __forceinline__ uint32_t alpha_bin_margin(uint32_t hit_ones, uint32_t entry, uint32_t &bin) {
    const uint32_t d = hit_ones - entry;
    const unsigned long long p = (unsigned long long)d * (unsigned long long)kNAngle;
    bin = (uint32_t)(p >> 32);
    return (uint32_t)p - kGuardLoA;
}

}  // namespace ppf
