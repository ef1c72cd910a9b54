// Step programs: a whole sampling timestep as ONE C call on caller-owned device memory.
//
// The per-timestep work of `Unet.forward` (src/DADiff.py:685-740) + `model_predictions` + the DDIM / posterior update
// (:1153-1209, 1221-1230, 1323-1344) is a fixed sequence of ~190 launches of this library's kernels over fixed buffers.  The
// host module that builds that sequence (founddiff_b200/engine.py) can RECORD it (founddiff_b200/program.py) into a plan
// file: the list of launches with their scalar arguments, every pointer argument as (allocation, byte offset), the packed
// weights' bytes, and the names of the buffers a caller feeds per step.  This file is the loader / executor:
//
//   fd_program_arena_bytes(path)                 how much device memory the plan needs (weights + activations + scratch)
//   fd_program_load(path, arena, bytes, &prog)   copies the weights into the caller's arena, resolves every pointer, builds
//                                                the TMA descriptors / GEMM plans of the convolutions (once)
//   fd_program_buffer(prog, "x_t", &bytes)       device address of a named buffer (x_t, x_input, time, coef, noise, ...)
//   fd_unet_step(prog, stream)                   launches the Unet evaluation of one timestep (conditioning + denoiser)
//   fd_sample_step(prog, stream)                 ... followed by the fused final_conv + model_predictions + update kernel
//
// No allocation, no synchronisation, no host reads inside a step: the calls are stream-ordered and CUDA-graph capturable.
// A C / C++ host therefore runs the sampler without Python or torch (examples/c_host.c, tests/test_gpu_program.py).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "fd_common.cuh"

struct fd_arg {
    void* p;
    long l;
    int i;
    float f;
};

#include "fd_program_dispatch.inc"

namespace {

constexpr char kMagic[8] = {'F', 'D', 'P', 'R', 'O', 'G', '2', 0};
constexpr int kMaxArgs = 40;
enum ArgKind : uint32_t { A_INT = 0, A_LONG = 1, A_FLOAT = 2, A_PTR = 3, A_NULL = 4, A_STREAM = 5 };
enum OpKind : uint32_t { OP_CALL = 0, OP_CONV = 1, OP_MEMSET = 2 };

struct FileHeader {
    char magic[8];
    uint32_t n_alloc, n_ops, n_unet_ops, reserved;
    uint64_t arena_bytes, data_bytes;
};
struct FileAlloc {
    uint64_t nbytes, offset, data_offset;      // offset: into the arena; data_offset: into the file's data blob (has_data)
    uint32_t has_data, reserved;
    char name[48];
};
struct FileArg {
    uint32_t kind, alloc;
    uint64_t value;                            // int / long / float bits / byte offset inside the allocation
};
struct FileOp {
    uint32_t kind, nargs;
    char fn[40];
    FileArg args[kMaxArgs];
};

struct Op {
    uint32_t kind;
    int fn;
    fd_arg a[kMaxArgs];
    fd_conv_params conv;
    fd_gemm_plan* plan;
    size_t memset_bytes;
};

}  // namespace

struct fd_program {
    std::vector<FileAlloc> allocs;
    std::vector<Op> ops;
    uint32_t n_unet_ops;
    char* arena;
};

static bool read_exact(FILE* f, void* dst, size_t n) { return fread(dst, 1, n, f) == n; }

extern "C" long fd_program_arena_bytes(const char* path) {
    if (!path) return -1;
    FILE* f = fopen(path, "rb");
    if (!f) return -1;
    FileHeader h;
    const bool ok = read_exact(f, &h, sizeof(h)) && memcmp(h.magic, kMagic, 8) == 0;
    fclose(f);
    return ok ? (long)h.arena_bytes : -1;
}

extern "C" void fd_program_destroy(fd_program* prog) {
    if (!prog) return;
    for (Op& op : prog->ops)
        if (op.plan) fd_conv2d_tc_plan_destroy(op.plan);
    delete prog;
}

// fd_conv_params as the recorder serialises it: 12 pointers (src0, src1, weight, bias, gate, addend, out, gn_sums, weight_up4,
// gn_ws, ln_v, ln_rstd), then the integer fields in struct order, ln_eps as the only float
static int unpack_conv(const FileOp& fo, const std::vector<void*>& ptr, fd_conv_params* p) {
    if (fo.nargs != 12 + 20) return FD_ERR_BAD_ARGUMENT;
    memset(p, 0, sizeof(*p));
    const void** pp[12] = {&p->src0, &p->src1, &p->weight, (const void**)&p->bias, (const void**)&p->gate, &p->addend,
                           (const void**)&p->out, (const void**)&p->gn_sums, &p->weight_up4, (const void**)&p->gn_ws,
                           (const void**)&p->ln_v, (const void**)&p->ln_rstd};
    for (int i = 0; i < 12; ++i) *pp[i] = ptr[i];
    int* ip[19] = {&p->c0, &p->c1, &p->ld0, &p->B, &p->Hin, &p->Win, &p->Cout, &p->KH, &p->KW, &p->stride, &p->pad, &p->upsample,
                   &p->silu_from, &p->gate_stride, &p->gn_groups, &p->per_batch_weight, &p->dtype, &p->relu_out, &p->ab_dtype_p1};
    for (int i = 0; i < 19; ++i) *ip[i] = (int)(int64_t)fo.args[12 + i].value;
    uint32_t bits = (uint32_t)fo.args[12 + 19].value;
    memcpy(&p->ln_eps, &bits, 4);
    return 0;
}

extern "C" int fd_program_load(const char* path, void* arena, long arena_bytes, fd_program** out) {
    if (!path || !arena || !out) return FD_ERR_BAD_ARGUMENT;
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) return FD_ERR_BAD_ARGUMENT;
    FileHeader h;
    if (!read_exact(f, &h, sizeof(h)) || memcmp(h.magic, kMagic, 8) != 0 || (long)h.arena_bytes > arena_bytes) { fclose(f); return FD_ERR_BAD_ARGUMENT; }
    fd_program* prog = new fd_program();
    prog->arena = (char*)arena;
    prog->n_unet_ops = h.n_unet_ops;
    prog->allocs.resize(h.n_alloc);
    std::vector<FileOp> fops(h.n_ops);
    bool ok = read_exact(f, prog->allocs.data(), sizeof(FileAlloc) * h.n_alloc) && read_exact(f, fops.data(), sizeof(FileOp) * h.n_ops);
    const long data_start = ok ? ftell(f) : 0;
    int rc = ok ? 0 : FD_ERR_BAD_ARGUMENT;
    // allocations: constants / initial contents from the file, everything else zero
    std::vector<char> host;
    for (uint32_t i = 0; rc == 0 && i < h.n_alloc; ++i) {
        const FileAlloc& al = prog->allocs[i];
        if (al.offset + al.nbytes > h.arena_bytes) { rc = FD_ERR_BAD_ARGUMENT; break; }
        if (al.has_data) {
            host.resize(al.nbytes);
            if (fseek(f, data_start + (long)al.data_offset, SEEK_SET) != 0 || !read_exact(f, host.data(), al.nbytes)) { rc = FD_ERR_BAD_ARGUMENT; break; }
            if (cudaMemcpy(prog->arena + al.offset, host.data(), al.nbytes, cudaMemcpyHostToDevice) != cudaSuccess) rc = (int)cudaGetLastError();
        } else if (cudaMemset(prog->arena + al.offset, 0, al.nbytes) != cudaSuccess) {
            rc = (int)cudaGetLastError();
        }
    }
    fclose(f);
    // operations: resolve pointers, look up entry points, build convolution plans
    prog->ops.resize(rc == 0 ? h.n_ops : 0);
    for (uint32_t i = 0; rc == 0 && i < h.n_ops; ++i) {
        const FileOp& fo = fops[i];
        Op& op = prog->ops[i];
        memset(&op, 0, sizeof(op));
        op.kind = fo.kind;
        if (fo.nargs > (uint32_t)kMaxArgs) { rc = FD_ERR_BAD_ARGUMENT; break; }
        std::vector<void*> ptr(fo.nargs, nullptr);
        for (uint32_t j = 0; j < fo.nargs; ++j) {
            const FileArg& fa = fo.args[j];
            fd_arg& a = op.a[j];
            if (fa.kind == A_PTR) {
                if (fa.alloc >= h.n_alloc || fa.value > prog->allocs[fa.alloc].nbytes) { rc = FD_ERR_BAD_ARGUMENT; break; }
                a.p = ptr[j] = prog->arena + prog->allocs[fa.alloc].offset + fa.value;
            } else if (fa.kind == A_INT) {
                a.i = (int)(int64_t)fa.value;
            } else if (fa.kind == A_LONG) {
                a.l = (long)(int64_t)fa.value;
            } else if (fa.kind == A_FLOAT) {
                const uint32_t bits = (uint32_t)fa.value;
                memcpy(&a.f, &bits, 4);
            }
        }
        if (rc) break;
        if (fo.kind == OP_CALL) {
            op.fn = -1;
            for (int k = 0; k < kNumFns; ++k)
                if (strncmp(fo.fn, kFnNames[k], sizeof(fo.fn)) == 0) op.fn = k;
            if (op.fn < 0) rc = FD_ERR_UNSUPPORTED;
        } else if (fo.kind == OP_CONV) {
            rc = unpack_conv(fo, ptr, &op.conv);
            if (rc == 0 && fd_conv2d_tc_supported(&op.conv)) rc = fd_conv2d_tc_plan_create(&op.conv, &op.plan);
        } else if (fo.kind == OP_MEMSET) {
            op.memset_bytes = (size_t)fo.args[1].value;
        } else {
            rc = FD_ERR_BAD_ARGUMENT;
        }
    }
    if (rc) { fd_program_destroy(prog); return rc; }
    *out = prog;
    return 0;
}

extern "C" void* fd_program_buffer(const fd_program* prog, const char* name, long* nbytes) {
    if (!prog || !name) return nullptr;
    for (const FileAlloc& al : prog->allocs)
        if (strncmp(al.name, name, sizeof(al.name)) == 0) {
            if (nbytes) *nbytes = (long)al.nbytes;
            return prog->arena + al.offset;
        }
    return nullptr;
}

static int run_range(const fd_program* prog, uint32_t first, uint32_t last, cudaStream_t st) {
    if (!prog) return FD_ERR_BAD_ARGUMENT;
    for (uint32_t i = first; i < last && i < prog->ops.size(); ++i) {
        const Op& op = prog->ops[i];
        int rc = 0;
        if (op.kind == OP_CALL) rc = fd_dispatch(op.fn, op.a, st);
        else if (op.kind == OP_CONV) rc = op.plan ? fd_conv2d_tc_run(op.plan, st) : fd_conv2d_simt(&op.conv, st);
        else if (cudaMemsetAsync(op.a[0].p, 0, op.memset_bytes, st) != cudaSuccess) rc = (int)cudaGetLastError();
        if (rc) return rc;
    }
    return 0;
}

extern "C" int fd_unet_step(const fd_program* prog, cudaStream_t stream) { return run_range(prog, 0, prog ? prog->n_unet_ops : 0, stream); }
extern "C" int fd_sample_step(const fd_program* prog, cudaStream_t stream) { return run_range(prog, 0, prog ? (uint32_t)prog->ops.size() : 0, stream); }
extern "C" int fd_program_num_launches(const fd_program* prog) { return prog ? (int)prog->ops.size() : -1; }
