// kc_tape.h -- the op tape of the fused elementwise kernel: shared by the host planner
// (kc_fusion.cu), the ahead-of-time kernels (kc_kernels.cu) and the kernels specialised at run
// time with NVRTC (kc_jit.cu embeds this file and kc_tile_vm.cuh as their prelude).
#pragma once
#ifdef __CUDACC_RTC__
typedef unsigned char uint8_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#else
#include <cstdint>
#endif

// ---- the fused elementwise kernel's op tape ---------------------------------
// An accumulator machine interpreted once per float4 of pixels, everything in
// registers: S[k] are the float4s loaded from the kernel's source planes, T[j]
// temporaries, `acc` the accumulator.  One instruction = op | (arg << 8).
// arg addresses an operand: 0..7 = S[arg], 8..13 = T[arg-8], 14 = immediate.
constexpr int KC_MAX_SRC = 8;
constexpr int KC_SRC_SOFT_CAP = 8;   // default bound for merging outputs that only share sources (kc_fusion.cu)
constexpr int KC_MAX_TMP = 6;
constexpr int KC_MAX_OUT = 6;
constexpr int KC_MAX_TAPE = 96;   // instructions per launch, all segments together
constexpr int KC_MAX_SEG = 4;     // independent programs per launch (blockIdx.y)
constexpr int KC_ARG_TMP0 = 8;
constexpr int KC_ARG_IMM = 14;

enum KcTapeOp : uint32_t {
    TOP_LD = 0,      // acc = X
    TOP_ADD,         // acc = acc + X
    TOP_SUB,         // acc = acc - X
    TOP_RSUB,        // acc = X - acc
    TOP_MUL,         // acc = acc * X
    TOP_DIV,         // acc = acc / X
    TOP_RDIV,        // acc = X / acc
    TOP_POW,         // acc = pow(acc, X)
    TOP_RPOW,        // acc = pow(X, acc)
    TOP_ST_TMP,      // T[arg] = acc
    TOP_ST_OUT,      // out[arg][i] = acc
    TOP_PACK_RGBA,   // rgba8[i] = to_u8(T0,T1,T2,acc)   (arg = 1: sRGB transfer on T0..T2)
    TOP_PACK_GRAY,   // rgba8[i] = (to_u8(acc) x3, 255)  (arg = 1: sRGB)
};

// One launch runs up to KC_MAX_SEG independent segments (blockIdx.y picks one):
// each has its own source/output planes and its own slice of the tape.  The
// three channels of an Rgba Mix chain are three segments of one launch.
struct KcSegment {
    const float* src[KC_MAX_SRC];
    float* out[KC_MAX_OUT];
    uint32_t* out_rgba8;
    uint32_t tape_begin, tape_end;
    uint32_t n_src, pad;
};

struct KcTapeArgs {
    KcSegment seg[KC_MAX_SEG];
    unsigned long long n;  // pixels per plane (the same for every segment of a launch)
    uint32_t n_seg;
    uint32_t variant;      // number of shared-memory temporaries the tapes of this launch touch
    uint32_t instr[KC_MAX_TAPE];
    float imm[KC_MAX_TAPE];
};

