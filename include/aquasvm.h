/* aquasvm.h -- the scalar programs of device-side loops.
 *
 * The reference evaluates the expressions of its set_scalar / if / while / assert tools on
 * the host, behind the events of the variables they read (SetScalar.cpp:146-195,
 * Conditional.cpp:85-96): every reduction inside a `while` costs a device -> host round
 * trip before the host knows what to enqueue next.  Here a loop whose body only holds
 * capturable tools runs as ONE CUDA graph with a WHILE conditional node (csrc/devloop.cu);
 * the scalar tools of the body become small stack programs over a TABLE of typed scalars
 * resident in device memory, interpreted by one thread.
 *
 * The program format is part of the C-ABI (aqc_loop_svm): the host (host/devloop.cpp)
 * compiles the expression grammar of host/tokenizer.hpp into it.  The interpreter below is
 * shared source: __device__ inside libaquacuda.so, plain C++ for the CPU tests that pin it
 * to the host evaluator (tests/test_host_cpu.py).  Arithmetic is IEEE double, one operation
 * at a time, narrowed once per store exactly like Variables::solve (host/variables.cpp) --
 * + - * / and the comparisons give the host's bits; the libm functions can differ from
 * glibc's in the last place.
 */
#ifndef AQUASVM_H
#define AQUASVM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    AQS_IMM = 0,   /* push imm */
    AQS_LOAD,      /* push table[a] read as kind b ('f', 'u', 'i') */
    AQS_STORE,     /* pop -> table[a] narrowed to kind b (range-checked like narrow_cast) */
    AQS_ADD, AQS_SUB, AQS_MUL, AQS_DIV, AQS_MOD, AQS_POW,
    AQS_NEG, AQS_NOT,
    AQS_LT, AQS_GT, AQS_LE, AQS_GE, AQS_EQ, AQS_NE, AQS_AND, AQS_OR,
    AQS_SELECT,    /* pop b, a, c; push c != 0 ? a : b (both branches evaluated, like the host) */
    AQS_CALL,      /* a = function (AQS_F_*), b = number of arguments */
    AQS_FOLD,      /* reduction epilogue: table[a] = op(table[b], table[c]) per component;
                      imm packs op (AQC_OP_*), kind and ncomp: op + 4 * kindcode + 16 * ncomp */
    AQS_SNAP,      /* report tool `a`: append the table to the history */
    AQS_ASSERT,    /* pop, narrowed to int like the host's solve("int"); zero -> error code 0x10000 + a */
    AQS_SETCOND,   /* pop, narrowed to int -> loop condition; a true one counts one iteration (the
                      body runs next), and ends the loop with error 0x30000 at max_iters */
    AQS_RECOND     /* loop condition = the one the last AQS_SETCOND left in the header (nothing is
                      counted): the entry program of a loop whose first pass ran before the graph */
};

enum {
    AQS_F_SQRT = 0, AQS_F_ABS, AQS_F_SIN, AQS_F_COS, AQS_F_TAN, AQS_F_ASIN, AQS_F_ACOS,
    AQS_F_ATAN, AQS_F_ATAN2, AQS_F_SINH, AQS_F_COSH, AQS_F_TANH, AQS_F_EXP, AQS_F_LOG,
    AQS_F_LOG2, AQS_F_LOG10, AQS_F_FLOOR, AQS_F_CEIL, AQS_F_RINT, AQS_F_SIGN,
    AQS_F_MIN, AQS_F_MAX, AQS_F_SUM, AQS_F_AVG
};

typedef struct aqs_op {
    int32_t code;
    int32_t a, b, c;
    double imm;
} aqs_op;

/* What a loop leaves behind for the host, in front of the table in the pinned mirror */
typedef struct aqs_header {
    uint32_t iters;  /* conditions that came out true = executions of the body */
    uint32_t snaps;  /* AQS_SNAP executed (the history keeps the first hist_rows) */
    uint32_t error;  /* 0, 1 + index of the op (within its program) whose store overflowed,
                        0x10000 + assert id, 0x20000: stack overflow, 0x30000: max_iters reached,
                        0x40000: the condition overflows an int */
    uint32_t cond;   /* last condition */
} aqs_header;

#define AQS_STACK 32

#ifdef __cplusplus
}

#if defined(__CUDACC__)
#define AQS_HD __host__ __device__ __forceinline__
#else
#include <cmath>
#define AQS_HD inline
#endif

/* (device: the _rn intrinsics keep nvcc from contracting a * b + c of two interpreted ops
 * that it may see back to back after unrolling) */
AQS_HD double aqs_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
AQS_HD double aqs_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

AQS_HD double aqs_load(const char* tab, int off, int kind)
{
    switch (kind) {
        case 'i': return (double)*(const int32_t*)(tab + off);
        case 'u': return (double)*(const uint32_t*)(tab + off);
        default: return (double)*(const float*)(tab + off);
    }
}

/* narrow_cast (host/variables.cpp store): truncation with a range check; false on overflow */
AQS_HD bool aqs_store(char* tab, int off, int kind, double v)
{
    switch (kind) {
        case 'i':
            if (!(v >= -2147483648.0 && v <= 2147483647.0))
                return false;
            *(int32_t*)(tab + off) = (int32_t)v;
            return true;
        case 'u':
            if (!(v >= 0.0 && v <= 4294967295.0))
                return false;
            *(uint32_t*)(tab + off) = (uint32_t)v;
            return true;
        default:
            *(float*)(tab + off) = (float)v;
            return true;
    }
}

AQS_HD double aqs_call(int f, const double* a, int n)
{
    switch (f) {
        case AQS_F_SQRT: return sqrt(a[0]);
        case AQS_F_ABS: return fabs(a[0]);
        case AQS_F_SIN: return sin(a[0]);
        case AQS_F_COS: return cos(a[0]);
        case AQS_F_TAN: return tan(a[0]);
        case AQS_F_ASIN: return asin(a[0]);
        case AQS_F_ACOS: return acos(a[0]);
        case AQS_F_ATAN: return atan(a[0]);
        case AQS_F_ATAN2: return atan2(a[0], a[1]);
        case AQS_F_SINH: return sinh(a[0]);
        case AQS_F_COSH: return cosh(a[0]);
        case AQS_F_TANH: return tanh(a[0]);
        case AQS_F_EXP: return exp(a[0]);
        case AQS_F_LOG: return log(a[0]);
        case AQS_F_LOG2: return log2(a[0]);
        case AQS_F_LOG10: return log10(a[0]);
        case AQS_F_FLOOR: return floor(a[0]);
        case AQS_F_CEIL: return ceil(a[0]);
        case AQS_F_RINT: return rint(a[0]);
        case AQS_F_SIGN: return (double)((a[0] > 0) - (a[0] < 0));
        default: {
            double r = a[0];
            for (int k = 1; k < n; k++)
                r = f == AQS_F_MIN ? fmin(r, a[k]) : (f == AQS_F_MAX ? fmax(r, a[k]) : aqs_add(r, a[k]));
            return f == AQS_F_AVG ? r / n : r;
        }
    }
}

/* The fold of a reduction's raw result with the tool's null value (Reduction.hcl.in:90;
 * host: Reduction::_execute), in the array's own arithmetic */
AQS_HD void aqs_fold(char* tab, int dst, int raw, int ident, int packed)
{
    const int op = packed & 3, kc = (packed >> 2) & 3, n = packed >> 4;
    for (int c = 0; c < n; c++) {
        if (kc == 0) {
            const float a = *(const float*)(tab + raw + 4 * c), b = *(const float*)(tab + ident + 4 * c);
#if defined(__CUDA_ARCH__)
            const float r = op == 0 ? __fadd_rn(a, b) : (op == 1 ? fminf(a, b) : fmaxf(a, b));
#else
            const float r = op == 0 ? a + b : (op == 1 ? fminf(a, b) : fmaxf(a, b));
#endif
            *(float*)(tab + dst + 4 * c) = r;
        } else if (kc == 1) {
            const uint32_t a = *(const uint32_t*)(tab + raw + 4 * c), b = *(const uint32_t*)(tab + ident + 4 * c);
            *(uint32_t*)(tab + dst + 4 * c) = op == 0 ? a + b : (op == 1 ? (a < b ? a : b) : (a > b ? a : b));
        } else {
            const int32_t a = *(const int32_t*)(tab + raw + 4 * c), b = *(const int32_t*)(tab + ident + 4 * c);
            *(int32_t*)(tab + dst + 4 * c) = op == 0 ? a + b : (op == 1 ? (a < b ? a : b) : (a > b ? a : b));
        }
    }
}

/* Runs `n` ops over the table.  hist: hist_rows rows of table_bytes each (may be NULL).
 * Returns the last condition set by AQS_SETCOND, or -1 when the program holds none. */
AQS_HD int aqs_run(const aqs_op* prog, int n, char* tab, int table_bytes, aqs_header* hdr, char* hist,
                   int hist_rows, uint32_t max_iters)
{
    double st[AQS_STACK];
    int sp = 0, cond = -1;
    for (int k = 0; k < n; k++) {
        const aqs_op op = prog[k];
        switch (op.code) {
            case AQS_IMM:
                if (sp >= AQS_STACK) { hdr->error = 0x20000; return cond; }
                st[sp++] = op.imm;
                break;
            case AQS_LOAD:
                if (sp >= AQS_STACK) { hdr->error = 0x20000; return cond; }
                st[sp++] = aqs_load(tab, op.a, op.b);
                break;
            case AQS_STORE:
                if (!aqs_store(tab, op.a, op.b, st[--sp]) && !hdr->error)
                    hdr->error = 1u + (uint32_t)k;
                break;
            case AQS_ADD: sp--; st[sp - 1] = aqs_add(st[sp - 1], st[sp]); break;
            case AQS_SUB: sp--; st[sp - 1] = aqs_add(st[sp - 1], -st[sp]); break;
            case AQS_MUL: sp--; st[sp - 1] = aqs_mul(st[sp - 1], st[sp]); break;
            case AQS_DIV: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
            case AQS_MOD: sp--; st[sp - 1] = fmod(st[sp - 1], st[sp]); break;
            case AQS_POW: sp--; st[sp - 1] = pow(st[sp - 1], st[sp]); break;
            case AQS_NEG: st[sp - 1] = -st[sp - 1]; break;
            case AQS_NOT: st[sp - 1] = st[sp - 1] == 0.0 ? 1.0 : 0.0; break;
            case AQS_LT: sp--; st[sp - 1] = st[sp - 1] < st[sp] ? 1.0 : 0.0; break;
            case AQS_GT: sp--; st[sp - 1] = st[sp - 1] > st[sp] ? 1.0 : 0.0; break;
            case AQS_LE: sp--; st[sp - 1] = st[sp - 1] <= st[sp] ? 1.0 : 0.0; break;
            case AQS_GE: sp--; st[sp - 1] = st[sp - 1] >= st[sp] ? 1.0 : 0.0; break;
            case AQS_EQ: sp--; st[sp - 1] = st[sp - 1] == st[sp] ? 1.0 : 0.0; break;
            case AQS_NE: sp--; st[sp - 1] = st[sp - 1] != st[sp] ? 1.0 : 0.0; break;
            case AQS_AND: sp--; st[sp - 1] = (st[sp - 1] != 0.0 && st[sp] != 0.0) ? 1.0 : 0.0; break;
            case AQS_OR: sp--; st[sp - 1] = (st[sp - 1] != 0.0 || st[sp] != 0.0) ? 1.0 : 0.0; break;
            case AQS_SELECT:
                sp -= 2;
                st[sp - 1] = st[sp - 1] != 0.0 ? st[sp] : st[sp + 1];
                break;
            case AQS_CALL:
                sp -= op.b;
                st[sp] = aqs_call(op.a, st + sp, op.b);
                sp++;
                break;
            case AQS_FOLD: aqs_fold(tab, op.a, op.b, op.c, (int)op.imm); break;
            case AQS_SNAP:
                if (hist && (int)hdr->snaps < hist_rows) {
                    char* row = hist + (size_t)hdr->snaps * (size_t)(table_bytes + 16);
                    *(int32_t*)row = op.a;
                    for (int b = 0; b < table_bytes; b += 4)
                        *(uint32_t*)(row + 16 + b) = *(const uint32_t*)(tab + b);
                }
                hdr->snaps++;
                break;
            case AQS_ASSERT: {
                const double v = st[--sp];
                if (!(v >= -2147483648.0 && v <= 2147483647.0)) {
                    if (!hdr->error)
                        hdr->error = 0x40000u;
                } else if ((int32_t)v == 0 && !hdr->error)
                    hdr->error = 0x10000u + (uint32_t)op.a;
                break;
            }
            case AQS_SETCOND: {
                const double v = st[--sp];
                if (!(v >= -2147483648.0 && v <= 2147483647.0)) {
                    if (!hdr->error)
                        hdr->error = 0x40000u;
                    cond = 0;
                } else
                    cond = (int32_t)v != 0 ? 1 : 0;
                if (hdr->error)
                    cond = 0; /* the host would have thrown where the error arose */
                if (cond) {
                    if (hdr->iters >= max_iters) {
                        hdr->error = 0x30000u;
                        cond = 0;
                    } else
                        hdr->iters++;
                }
                hdr->cond = (uint32_t)cond;
                break;
            }
            case AQS_RECOND: cond = (hdr->cond != 0 && !hdr->error) ? 1 : 0; break;
            default: break;
        }
    }
    return cond;
}
#endif /* __cplusplus */

#endif /* AQUASVM_H */
