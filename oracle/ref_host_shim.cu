// oracle/ref_host_shim.cu -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-callable shims around the HOST compilation of the reference's own
// __host__ __device__ functions (kernel.cu is compiled in place from
// /root/reference, see oracle/Makefile). They let the CPU test-suite pin
// oracle/ppf_oracle.c against the reference's own code without a GPU.
// Glue only: no arithmetic is done here.
#include <cstring>
#include <vector_types.h>
#include "kernel.h"        // reference header (unmodified)
#include "vector_ops.h"    // reference header (unmodified)

static float3 f3(const float *p) { return make_float3(p[0], p[1], p[2]); }

extern "C" {

unsigned int refhost_hash(const void *bytes, int n) { return hash((void *)bytes, n); }

void refhost_compute_ppf(const float *p1, const float *n1, const float *p2, const float *n2, float *out4) {
    float4 f = compute_ppf(f3(p1), f3(n1), f3(p2), f3(n2));
    out4[0] = f.x; out4[1] = f.y; out4[2] = f.z; out4[3] = f.w;
}

void refhost_disc_feature(const float *in4, float d_dist, float d_angle, float *out4) {
    float4 f = disc_feature(make_float4(in4[0], in4[1], in4[2], in4[3]), d_dist, d_angle);
    out4[0] = f.x; out4[1] = f.y; out4[2] = f.z; out4[3] = f.w;
}

float refhost_quant_downf(float x, float y) { return quant_downf(x, y); }
float refhost_d_angle0(void) { return D_ANGLE0; }

void refhost_discretize(const float *in3, float d, float *out3) {
    float3 r = discretize(f3(in3), d);
    out3[0] = r.x; out3[1] = r.y; out3[2] = r.z;
}

void refhost_rot(int axis, float theta, float *T16) {
    float T[4][4];
    if (axis == 0) rotx(theta, T); else if (axis == 1) roty(theta, T); else rotz(theta, T);
    std::memcpy(T16, T, sizeof(T));
}
void refhost_trans(const float *v3, float *T16) {
    float T[4][4];
    trans(f3(v3), T);
    std::memcpy(T16, T, sizeof(T));
}
void refhost_mat4f_mul(const float *A16, const float *B16, float *C16) {
    float C[4][4];
    mat4f_mul((const float (*)[4])A16, (const float (*)[4])B16, C);
    std::memcpy(C16, C, sizeof(C));
}
void refhost_mat4f_vmul(const float *A16, const float *b4, float *c4) {
    float4 c = mat4f_vmul((const float (*)[4])A16, make_float4(b4[0], b4[1], b4[2], b4[3]));
    c4[0] = c.x; c4[1] = c.y; c4[2] = c.z; c4[3] = c.w;
}
void refhost_invht(const float *T16, float *Tinv16) {
    float T[4][4], Ti[4][4];
    std::memcpy(T, T16, sizeof(T));
    invht(T, Ti);
    std::memcpy(Tinv16, Ti, sizeof(Ti));
}
void refhost_hrotmat2quat(const float *T16, float *q4) {
    float T[4][4];
    std::memcpy(T, T16, sizeof(T));
    float4 q = hrotmat2quat(T);
    q4[0] = q.x; q4[1] = q.y; q4[2] = q.z; q4[3] = q.w;
}
void refhost_cross(const float *u, const float *v, float *w) {
    float3 r = cross(f3(u), f3(v));
    w[0] = r.x; w[1] = r.y; w[2] = r.z;
}
float refhost_dot3(const float *u, const float *v) { return dot(f3(u), f3(v)); }

}  // extern "C"
