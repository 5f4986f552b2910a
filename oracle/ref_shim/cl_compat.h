/*
 * cl_compat.h -- the OpenCL C surface that the reference's kernel text
 * (src/intersect_kernel.cl) needs in order to be compiled, unmodified apart
 * from the two cast-syntax rewrites done by oracle/Makefile, as a C++
 * translation unit for the CPU.  TEST INFRASTRUCTURE ONLY.
 *
 * The arithmetic of every built-in is pinned to: binary32, one rounding per
 * operation, no contraction (the TU is built with -ffp-contract=off), dot
 * summed ((x+y)+z)+w, cross with w = 0, normalize(v) = v / sqrt(dot(v,v)),
 * max(a,b) = a < b ? b : a, min(a,b) = b < a ? b : a,
 * clamp(v,lo,hi) = fmin(fmax(v,lo),hi).  (OpenCL 1.2 spec, 6.12.2/6.12.4/6.12.5.)
 */
#pragma once
#include <cmath>
#include <cstdint>

typedef unsigned int uint;
#define __global
#define __kernel

struct float4 {
	float x, y, z, w;
	float4() {}
	explicit float4(float s) : x(s), y(s), z(s), w(s) {}
	float4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
};
struct uint4 { uint x, y, z, w; };
struct int2 { int x, y; };

/* `(float4) (a,b,c,d)` / `(float4) (s)` / `(int2) (a,b)` vector literals */
static inline float4 mk4(float a, float b, float c, float d) { return float4(a, b, c, d); }
static inline float4 mk4(float s) { return float4(s); }
static inline int2 mk2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }

static inline float4 operator+(float4 a, float4 b) { return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
static inline float4 operator-(float4 a, float4 b) { return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
static inline float4 operator*(float4 a, float4 b) { return float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
static inline float4 operator*(float s, float4 a) { return float4(s * a.x, s * a.y, s * a.z, s * a.w); }
static inline float4 operator*(float4 a, float s) { return float4(a.x * s, a.y * s, a.z * s, a.w * s); }

static inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
static inline float4 cross(float4 a, float4 b)
{
	return float4(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x, 0.0f);
}
static inline float length(float4 a) { return std::sqrt(dot(a, a)); }
static inline float4 normalize(float4 a)
{
	const float l = length(a);
	return float4(a.x / l, a.y / l, a.z / l, a.w / l);
}
static inline float clamp(float v, float lo, float hi) { return std::fmin(std::fmax(v, lo), hi); }
static inline float max(float a, float b) { return a < b ? b : a; }
static inline float min(float a, float b) { return b < a ? b : a; }
static inline int max(int a, int b) { return a < b ? b : a; }
/* Transcendentals (only the ambient-occlusion samplers use them).  OpenCL leaves their last bits to the
 * device (<= 4 ulp), so any definition is a legal reference; this one is chosen so that a CPU and a GPU can
 * agree bit for bit: evaluate in double precision and round once to float (two double-precision libraries
 * differ by far less than a float rounding step almost everywhere). */
static inline float sin(float x) { return (float)std::sin((double)x); }
static inline float cos(float x) { return (float)std::cos((double)x); }
static inline float acos(float x) { return (float)std::acos((double)x); }
static inline float cospi(float x) { return (float)std::cos(3.14159265358979323846 * (double)x); }
static inline float sinpi(float x) { return (float)std::sin(3.14159265358979323846 * (double)x); }
using std::fabs;
using std::sqrt;

/* work-item id of the "NDRange" the glue iterates */
extern thread_local uint cl_compat_gid[2];
static inline uint get_global_id(int d) { return cl_compat_gid[d]; }
