// jit.cuh -- late (plan-time) specialisation of the template kernel.
//
// The static template kernel (fastplan.cuh: tp_gather_kernel) treats accumulator positions as template DATA, so
// every contribution is a shared-memory read-modify-write.  Once the templates of a pattern are known they are
// constants: this file writes the same computation as straight-line CUDA source per template -- cell offsets, local
// indices and accumulator positions are literals, the accumulators are named scalars (registers), the closed-form
// entries of EvalBary are emitted term by term exactly as fp_bary_acc computes them -- compiles it with NVRTC for
// sm_100a and launches it on the SAME plan arrays (wdesc / slotpb / slotptr, same CTA packing, same write-out).
// Shared memory is used only for the transposing write-out.  Two kernels per plan (templates with short / long
// columns) keep the register allocation of the many short-column warps independent of the vertex columns.
// Any failure (no libnvrtc, compile error) silently keeps the static kernel; "template_jit" = 0 switches it off.
#pragma once
#include <dlfcn.h>
#include <nvrtc.h>

#include <sstream>

#include "common.cuh"
#include "fastplan.cuh"

namespace extfem {

constexpr int JIT_SPLIT_L = 40;   // columns longer than this run in the second (register-hungry) kernel

struct NvrtcApi {
    void *handle = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    bool load()
    {
        if (handle) return true;
        for (const char *n : {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"}) {
            handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (handle) break;
        }
        if (!handle) return false;
#define EXTFEM_NVRTC_SYM(name)                                   \
        *(void **)(&name) = dlsym(handle, "nvrtc" #name);        \
        if (!name) { handle = nullptr; return false; }
        EXTFEM_NVRTC_SYM(CreateProgram) EXTFEM_NVRTC_SYM(CompileProgram) EXTFEM_NVRTC_SYM(GetCUBINSize) EXTFEM_NVRTC_SYM(GetCUBIN)
        EXTFEM_NVRTC_SYM(GetProgramLogSize) EXTFEM_NVRTC_SYM(GetProgramLog) EXTFEM_NVRTC_SYM(DestroyProgram)
#undef EXTFEM_NVRTC_SYM
        return true;
    }
};
static NvrtcApi g_nvrtc;

// one template on the host: rounds as the device template words (fastplan.cuh: tp_tmpl_kernel)
struct JitTemplate {
    int r0 = 0, m = 0, L = 0;
    std::vector<unsigned> words;   // [m][TP_TW]
};

struct JitArgs {                   // mirrored in the generated source
    const int4 *wdesc;
    const int *slotpb;
    double *const *slotptr;
    const double *geo;
    long long Npad;
    int overwrite;
    int nlive;                     // warps of this launch
    const int *live;               // launch-order warp of every warp of this launch (full CTAs of one column-length class)
};

struct JitModule {
    bool tried = false, ok = false;
    cudaLibrary_t lib = nullptr;
    cudaKernel_t kernel[2] = {nullptr, nullptr};   // short / long columns
    bool present[2] = {false, false};
    std::string log;
    double compile_s = 0;
    ~JitModule() { if (lib) cudaLibraryUnload(lib); }
};

// ---- source generation ---------------------------------------------------------------------------------
struct JitGen {
    int dim, order, nv, ns;
    static int edge_a(int dim, int e) { return dim == 1 ? 0 : (dim == 2 ? e : (e < 3 ? 0 : (e < 5 ? 1 : 2))); }
    static int edge_b(int dim, int e) { return dim == 1 ? 1 : (dim == 2 ? (e + 1) % 3 : (e < 3 ? e + 1 : (e < 5 ? e - 1 : 3))); }
    int pair_index(int a, int b) const { return a * (2 * dim + 1 - a) / 2 + (b - a - 1); }
    // symbol of M[x][y]; records what the round needs
    std::string M(int x, int y, unsigned &planes, unsigned &diags, const std::string &R) const
    {
        if (x == y) { diags |= 1u << x; return R + "d" + std::to_string(x); }
        const int p = pair_index(std::min(x, y), std::max(x, y));
        planes |= 1u << p;
        return R + "g" + std::to_string(p);
    }
    // statements that add A_loc[t][kl] to accumulator `acc` (mirrors fp_bary_acc term by term)
    void entry(std::ostringstream &o, const std::string &acc, int t, int kl, unsigned &planes, unsigned &diags, const std::string &R) const
    {
        auto fma = [&](double c, const std::string &m) { o << acc << " = fma(" << c << ".0, " << m << ", " << acc << "); "; };
        auto sub = [&](const std::string &m) { o << acc << " -= " << m << "; "; };
        if (order == 1) { o << acc << " += " << M(t, kl, planes, diags, R) << "; "; return; }
        const bool tv = t < nv, kv = kl < nv;
        if (tv && kv) {
            if (t == kl) fma(3, M(t, t, planes, diags, R)); else sub(M(t, kl, planes, diags, R));
            return;
        }
        if (tv != kv) {
            const int i = tv ? t : kl, e = (tv ? kl : t) - nv, a = edge_a(dim, e), b = edge_b(dim, e);
            if (dim == 3) {
                if (i == a) fma(3, M(i, b, planes, diags, R)); else sub(M(i, b, planes, diags, R));
                if (i == b) fma(3, M(i, a, planes, diags, R)); else sub(M(i, a, planes, diags, R));
            } else {
                if (i == a) fma(4, M(i, b, planes, diags, R));
                if (i == b) fma(4, M(i, a, planes, diags, R));
            }
            return;
        }
        const int a = edge_a(dim, t - nv), b = edge_b(dim, t - nv), c = edge_a(dim, kl - nv), d = edge_b(dim, kl - nv);
        fma(a == c ? 8 : 4, M(b, d, planes, diags, R));
        fma(a == d ? 8 : 4, M(b, c, planes, diags, R));
        fma(b == c ? 8 : 4, M(a, d, planes, diags, R));
        fma(b == d ? 8 : 4, M(a, c, planes, diags, R));
    }
    void template_function(std::ostringstream &o, const JitTemplate &T, int id) const
    {
        o << "static __device__ __forceinline__ void tmpl_" << id
          << "(const double* __restrict__ geo, long long Npad, int pb, double* accw, double* const* ptrs, double* gptr, int dw, int lane, int overwrite) {\n";
        o << "  double ";
        for (int p = 0; p < T.L; ++p) o << (p ? ", " : "") << "a" << p << " = 0.0";
        o << ";\n";
        // per round: block A = loads of the geometry planes the round reads (named per round), block B = the entries.
        // Emission order A0 A1 B0 A2 B1 ...: the loads of the next round are in flight during the current one.
        std::vector<std::string> A(T.m), B(T.m);
        for (int r = 0; r < T.m; ++r) {
            const unsigned *w = &T.words[(size_t)r * TP_TW];
            const int pd = (int)w[0], kl = (int)(w[1] & 0xff);
            const std::string R = "r" + std::to_string(r) + "_";
            std::ostringstream body;
            unsigned planes = 0, diags = 0;
            for (int t = 0; t < ns; ++t) {
                const int pos = (int)(w[2 + t] / (TP_LD * 8));
                entry(body, "a" + std::to_string(pos), t, kl, planes, diags, R);
                body << "\n    ";
            }
            // diagonals need every plane that touches their vertex
            for (int x = 0; x <= dim; ++x)
                if (diags >> x & 1)
                    for (int y = 0; y <= dim; ++y)
                        if (y != x) planes |= 1u << pair_index(std::min(x, y), std::max(x, y));
            std::ostringstream la, lb;
            la << "  const double* " << R << "p = geo + (pb + (" << pd << "));\n  ";
            for (int p = 0; p < dim * (dim + 1) / 2; ++p)
                if (planes >> p & 1) la << "const double " << R << "g" << p << " = __ldg(" << R << "p + " << p << " * Npad); ";
            la << "\n";
            lb << "  { ";
            for (int x = 0; x <= dim; ++x)
                if (diags >> x & 1) {
                    lb << "const double " << R << "d" << x << " = -(";
                    bool first = true;
                    for (int y = 0; y <= dim; ++y)
                        if (y != x) { lb << (first ? "" : " + ") << R << "g" << pair_index(std::min(x, y), std::max(x, y)); first = false; }
                    lb << "); ";
                }
            lb << "\n    " << body.str() << "}\n";
            A[r] = la.str(); B[r] = lb.str();
        }
        o << A[0];
        for (int r = 0; r < T.m; ++r) {
            if (r + 1 < T.m) o << A[r + 1];
            o << B[r];
        }
        // transposing write-out in slices of 32 positions through the warp's [32][33] shared tile
        for (int p0 = 0; p0 < T.L; p0 += 32) {
            const int cnt = std::min(32, T.L - p0);
            for (int p = 0; p < cnt; ++p) o << "  accw[" << p * TP_LD << "] = a" << p0 + p << ";\n";
            o << "  __syncwarp();\n  wout(accw - lane, ptrs, gptr, dw, lane, " << p0 << ", " << cnt << ", overwrite);\n  __syncwarp();\n";
        }
        o << "}\n";
    }
    std::string source(const std::vector<JitTemplate> &tmpls, int splitL, bool present[2]) const
    {
        std::ostringstream o;
        o << "struct JitArgs { const int4* wdesc; const int* slotpb; double* const* slotptr; const double* geo; long long Npad; int overwrite; int nlive; const int* live; };\n"
          << "// column j of the warp gets positions [p0, p0 + cnt) from the tile acc[position - p0][column] (leading dimension " << TP_LD << ")\n"
          << "static __device__ __noinline__ void wout(const double* acc, double* const* ptrs, double* gptr, int dw, int lane, int p0, int cnt, int overwrite) {\n"
          << "  const bool pin = lane < cnt;\n"
          << "  const double* src = acc + lane * " << TP_LD << ";\n"
          << "  if (dw) {\n"
          << "    double* q = (double*)__shfl_sync(0xffffffffu, (unsigned long long)gptr, 0) + p0 + lane;\n"
          << "    #pragma unroll 8\n"
          << "    for (int j = 0; j < 32; ++j) { double v = src[j]; double* qq = q + (long long)j * dw; if (pin) { if (!overwrite) v += *qq; __stcs(qq, v); } }\n"
          << "  } else {\n"
          << "    #pragma unroll 8\n"
          << "    for (int j = 0; j < 32; ++j) { double* qq = ptrs[j] + p0 + lane; double v = src[j]; if (pin) { if (!overwrite) v += *qq; __stcs(qq, v); } }\n"
          << "  }\n"
          << "}\n";
        for (size_t k = 0; k < tmpls.size(); ++k) template_function(o, tmpls[k], (int)k);
        present[0] = present[1] = false;
        for (int cls = 0; cls < 2; ++cls) {
            o << "extern \"C\" __global__ void __launch_bounds__(" << TP_MAXW * 32 << ", " << (cls == 0 ? 4 : 3) << ") tpj_" << cls << "(const JitArgs A) {\n"
              << "  extern __shared__ double sm[];\n"
              << "  const int lane = threadIdx.x & 31, wi = blockIdx.x * " << TP_MAXW << " + (threadIdx.x >> 5);\n"
              << "  if (wi >= A.nlive) return;\n"
              << "  const int wq = __ldg(A.live + wi);\n"
              << "  const int4 d = __ldg(A.wdesc + wq);\n"
              << "  double* const gptr = (double*)__ldg((const unsigned long long*)A.slotptr + (size_t)wq * 32 + lane);\n"
              << "  const int pb = __ldg(A.slotpb + (size_t)wq * 32 + lane);\n"
              << "  double* tile = sm + (threadIdx.x >> 5) * " << (32 * TP_LD + 32) << ";\n"
              << "  double** ptrs = (double**)(tile + " << 32 * TP_LD << ");\n"
              << "  ptrs[lane] = gptr;\n"
              << "  switch (d.x) {\n";
            for (size_t k = 0; k < tmpls.size(); ++k)
                if ((tmpls[k].L > splitL) == (cls == 1)) {
                    o << "    case " << tmpls[k].r0 << ": tmpl_" << k << "(A.geo, A.Npad, pb, tile + lane, ptrs, gptr, d.w, lane, A.overwrite); break;\n";
                    present[cls] = true;
                }
            o << "    default: break;\n  }\n}\n";
        }
        return o.str();
    }
};

// compile `src` for sm_100a; fills M.lib / kernels
static bool jit_compile(const std::string &src, const bool present[2], JitModule &M)
{
    if (!g_nvrtc.load()) { M.log = "libnvrtc.so.12 not found"; return false; }
    nvrtcProgram prog = nullptr;
    if (g_nvrtc.CreateProgram(&prog, src.c_str(), "extfem_templates.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { M.log = "nvrtcCreateProgram failed"; return false; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    const nvrtcResult r = g_nvrtc.CompileProgram(prog, 3, opts);
    size_t ls = 0;
    g_nvrtc.GetProgramLogSize(prog, &ls);
    if (ls > 1) { M.log.resize(ls); g_nvrtc.GetProgramLog(prog, &M.log[0]); }
    if (r != NVRTC_SUCCESS) { g_nvrtc.DestroyProgram(&prog); return false; }
    size_t cs = 0;
    g_nvrtc.GetCUBINSize(prog, &cs);
    std::vector<char> cubin(cs);
    g_nvrtc.GetCUBIN(prog, cubin.data());
    g_nvrtc.DestroyProgram(&prog);
    if (cudaLibraryLoadData(&M.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) { M.log += " cudaLibraryLoadData failed"; cudaGetLastError(); return false; }
    for (int c = 0; c < 2; ++c) {
        M.present[c] = present[c];
        if (!present[c]) continue;
        const std::string name = "tpj_" + std::to_string(c);
        if (cudaLibraryGetKernel(&M.kernel[c], M.lib, name.c_str()) != cudaSuccess) { M.log += " cudaLibraryGetKernel failed"; cudaGetLastError(); return false; }
    }
    return true;
}

} // namespace extfem
