// Model-specialised integrator kernels (host side): the tables of one model are written out as constants of a generated
// translation unit, compiled for sm_100a by NVRTC and loaded through the driver API.  Inside such a kernel every loop
// over the model tables has a compile-time trip count, every table entry (block kinds, storage slots, lever coefficients,
// inertial constants) folds into the instruction stream and every shared-memory address is `lane + constant`: the
// generic kernels spend about two thirds of their instructions fetching and indexing exactly that (kernels.cuh,
// setup_cta; dynamics.cuh, BIONC_TAB_LOOP).  NVRTC and the driver are opened with dlopen on first use: a library built
// and loaded without them keeps the generic kernels.
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "model.cuh"

namespace bionc {
namespace jit {

// the device sources (model.cuh, tmem.cuh, dynamics.cuh, kernels.cuh), embedded by build.py
#include "_embedded_sources.inc"

struct Api {
    void* nvrtc = nullptr;
    void* cuda = nullptr;
    bool nvrtc_ok = false, cuda_ok = false;
    decltype(&nvrtcCreateProgram) CreateProgram = nullptr;
    decltype(&nvrtcCompileProgram) CompileProgram = nullptr;
    decltype(&nvrtcDestroyProgram) DestroyProgram = nullptr;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize = nullptr;
    decltype(&nvrtcGetCUBIN) GetCUBIN = nullptr;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize = nullptr;
    decltype(&nvrtcGetProgramLog) GetProgramLog = nullptr;
    decltype(&nvrtcAddNameExpression) AddNameExpression = nullptr;
    decltype(&nvrtcGetLoweredName) GetLoweredName = nullptr;
    decltype(&nvrtcGetErrorString) GetErrorString = nullptr;
    CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
    CUresult (*GetErrorString_cu)(CUresult, const char**) = nullptr;
};

template <class F>
inline bool sym(void* lib, const char* name, F& f) {
    f = reinterpret_cast<F>(dlsym(lib, name));
    return f != nullptr;
}

inline Api& api() {
    static Api a;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"}) {
            if ((a.nvrtc = dlopen(n, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
        }
        if (a.nvrtc != nullptr) {
            a.nvrtc_ok = sym(a.nvrtc, "nvrtcCreateProgram", a.CreateProgram) && sym(a.nvrtc, "nvrtcCompileProgram", a.CompileProgram) &&
                         sym(a.nvrtc, "nvrtcDestroyProgram", a.DestroyProgram) && sym(a.nvrtc, "nvrtcGetCUBINSize", a.GetCUBINSize) &&
                         sym(a.nvrtc, "nvrtcGetCUBIN", a.GetCUBIN) && sym(a.nvrtc, "nvrtcGetProgramLogSize", a.GetProgramLogSize) &&
                         sym(a.nvrtc, "nvrtcGetProgramLog", a.GetProgramLog) && sym(a.nvrtc, "nvrtcAddNameExpression", a.AddNameExpression) &&
                         sym(a.nvrtc, "nvrtcGetLoweredName", a.GetLoweredName) && sym(a.nvrtc, "nvrtcGetErrorString", a.GetErrorString);
        }
        a.cuda = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (a.cuda != nullptr) {
            a.cuda_ok = sym(a.cuda, "cuModuleLoadData", a.ModuleLoadData) && sym(a.cuda, "cuModuleUnload", a.ModuleUnload) &&
                        sym(a.cuda, "cuModuleGetFunction", a.ModuleGetFunction) && sym(a.cuda, "cuFuncSetAttribute", a.FuncSetAttribute) &&
                        sym(a.cuda, "cuLaunchKernel", a.LaunchKernel) && sym(a.cuda, "cuGetErrorString", a.GetErrorString_cu);
        }
    });
    return a;
}

// ---- source generation ---------------------------------------------------------------------------------

// one struct as a flat aggregate initialiser; `schema` names the scalar members in order: d = double, i = int.  Doubles
// are written as hexadecimal floating literals (exact).
inline void emit_struct(std::string& out, const void* p, const char* schema) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    char buf[64];
    out += "{";
    size_t off = 0;
    for (const char* s = schema; *s; ++s) {
        if (*s == 'd') {
            off = (off + 7) & ~size_t(7);
            double v;
            std::memcpy(&v, b + off, 8);
            if (v != v) std::snprintf(buf, sizeof buf, "__builtin_nan(\"\")");
            else if (v - v != 0.0) std::snprintf(buf, sizeof buf, v > 0 ? "__builtin_huge_val()" : "-__builtin_huge_val()");
            else std::snprintf(buf, sizeof buf, "%a", v);
            off += 8;
        } else {
            int v;
            std::memcpy(&v, b + off, 4);
            std::snprintf(buf, sizeof buf, "%d", v);
            off += 4;
        }
        out += buf;
        out += s[1] ? "," : "";
    }
    out += "}";
}

inline std::string repeat(const char* unit, int n) {
    std::string s;
    for (int i = 0; i < n; ++i) s += unit;
    return s;
}

template <class T>
inline void emit_array(std::string& out, const char* type, const char* name, const std::vector<T>& v, const std::string& schema) {
    out += std::string("__device__ const ") + type + " " + name + "[" + std::to_string(v.empty() ? 1 : v.size()) + "] = {";
    if (v.empty()) {
        out += "{}";
    }
    for (size_t i = 0; i < v.size(); ++i) {
        emit_struct(out, &v[i], schema.c_str());
        out += i + 1 < v.size() ? ",\n  " : "";
    }
    out += "};\n";
}

// kernel variant of one call shape
struct Variant {
    bool general = false, stab = false, persys = false, forces = false;
    int maxt = 128, spc = 32;
    int ctas = 0;  // CTAs resident per SM for this call shape (launch plan): sizes the register budget of the build
    bool roll_solve = false;  // keep the loops of the joint-level factorisation (instruction footprint)
    bool compact = false;     // compact working set (serial joint-level schedule): no blob area, inverse pivots in place
    bool operator==(const Variant& o) const {
        return general == o.general && stab == o.stab && persys == o.persys && forces == o.forces && maxt == o.maxt && spc == o.spc;
    }
    std::string name() const {
        char buf[160];
        std::snprintf(buf, sizeof buf, "bionc::rk4_kernel<%s, %s, %d, %d, %d, %s>", general ? "true" : "false", stab ? "true" : "false", maxt, spc,
                      persys ? 1 : 0, forces ? "true" : "false");
        return buf;
    }
};

// the generated translation unit of one model (tables + the generic kernel sources) and the explicit instantiation of
// one integrator variant
inline std::string source(const DevHeader& hdr, const std::vector<DevSeg>& seg, const std::vector<DevBlk>& blk, const std::vector<DevSide>& side,
                          const std::vector<int>& symtab, const Variant& v) {
    static_assert(sizeof(DevHeader) % 4 == 0, "DevHeader: ints only");
    static_assert(sizeof(DevSeg) == 30 * 8 + 8 * 4, "DevSeg: update the schema below");
    static_assert(sizeof(DevBlk) == 16 * 8 + 8 * 4, "DevBlk: update the schema below");
    static_assert(sizeof(DevSide) == 8 * 4 + 4 * 8, "DevSide: update the schema below");
    std::string s;
    s += "// generated by bionc_cuda (jit.cuh): model tables as constants + the generic kernel sources\n";
    s += "#define BIONC_SPEC 1\n#define BIONC_DYNAMICS_ONLY 1\n";
    if (v.ctas > 0) s += "#define BIONC_SPEC_MINCTAS " + std::to_string(v.ctas) + "\n";
    if (v.roll_solve) s += "#define BIONC_SPEC_ROLL_SOLVE 1\n";
    if (v.compact) s += "#define BIONC_SPEC_COMPACT 1\n";
    s += "#include \"model.cuh\"\nnamespace bionc {\nnamespace spec {\n";
    s += "constexpr int kSegments = " + std::to_string(hdr.N) + ";\n";
    s += "__device__ const DevHeader hdr = ";
    emit_struct(s, &hdr, repeat("i", (int)(sizeof(DevHeader) / 4)).c_str());
    s += ";\n";
    emit_array(s, "DevSeg", "seg", seg, repeat("d", 30) + repeat("i", 8));
    emit_array(s, "DevBlk", "blk", blk, repeat("d", 16) + repeat("i", 8));
    emit_array(s, "DevSide", "side", side, repeat("i", 8) + repeat("d", 4));
    s += "__device__ const int sym[" + std::to_string(symtab.empty() ? 1 : symtab.size()) + "] = {";
    for (size_t i = 0; i < symtab.size(); ++i) s += std::to_string(symtab[i]) + (i + 1 < symtab.size() ? "," : "");
    if (symtab.empty()) s += "0";
    s += "};\n}  // namespace spec\n}  // namespace bionc\n#include \"kernels.cuh\"\n";
    s += "template __global__ void " + v.name() +
         "(const unsigned char*, int, bionc::FextArgs, double*, double*, const double*, bionc::DynOpts, double, const double*, long long, long long, "
         "int, double*, long long, int*, long long);\n";
    return s;
}

struct Compiled {
    Variant v;
    std::vector<char> cubin;
    std::string lowered;
    // per device: module and function
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
};

// NVRTC: source -> cubin for sm_100a.  Returns an empty string on success, else what went wrong (with the compiler's log).
inline std::string compile(const std::string& src, const Variant& v, Compiled& out, std::string* log_out = nullptr) {
    Api& a = api();
    if (!a.nvrtc_ok) return "libnvrtc not found (dlopen)";
    nvrtcProgram prog = nullptr;
    nvrtcResult r = a.CreateProgram(&prog, src.c_str(), "bionc_spec.cu", kEmbeddedCount, kEmbeddedSources, kEmbeddedNames);
    if (r != NVRTC_SUCCESS) return std::string("nvrtcCreateProgram: ") + a.GetErrorString(r);
    const std::string expr = v.name();
    a.AddNameExpression(prog, expr.c_str());
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "-default-device"};
    r = a.CompileProgram(prog, 4, opts);
    size_t log_size = 0;
    a.GetProgramLogSize(prog, &log_size);
    std::string log(log_size, '\0');
    if (log_size > 1) a.GetProgramLog(prog, &log[0]);
    if (log_out != nullptr) *log_out = log;
    if (r != NVRTC_SUCCESS) {
        a.DestroyProgram(&prog);
        return std::string("nvrtcCompileProgram: ") + a.GetErrorString(r) + "\n" + log;
    }
    const char* lowered = nullptr;
    r = a.GetLoweredName(prog, expr.c_str(), &lowered);
    if (r != NVRTC_SUCCESS || lowered == nullptr) {
        a.DestroyProgram(&prog);
        return std::string("nvrtcGetLoweredName: ") + a.GetErrorString(r);
    }
    out.lowered = lowered;
    size_t n = 0;
    a.GetCUBINSize(prog, &n);
    out.cubin.resize(n);
    a.GetCUBIN(prog, out.cubin.data());
    a.DestroyProgram(&prog);
    out.v = v;
    return n > 0 ? "" : "NVRTC produced no cubin";
}

inline std::string cu_error(CUresult r) {
    const char* s = nullptr;
    if (api().GetErrorString_cu != nullptr) api().GetErrorString_cu(r, &s);
    return s != nullptr ? s : "unknown driver error";
}

// load into the current context (the caller has made the model's device current through the runtime API)
inline std::string load(Compiled& c, int max_dynamic_smem) {
    Api& a = api();
    if (!a.cuda_ok) return "libcuda.so.1 not found (dlopen)";
    CUresult r = a.ModuleLoadData(&c.mod, c.cubin.data());
    if (r != CUDA_SUCCESS) return "cuModuleLoadData: " + cu_error(r);
    r = a.ModuleGetFunction(&c.fn, c.mod, c.lowered.c_str());
    if (r != CUDA_SUCCESS) return "cuModuleGetFunction(" + c.lowered + "): " + cu_error(r);
    r = a.FuncSetAttribute(c.fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, max_dynamic_smem);
    if (r != CUDA_SUCCESS) return "cuFuncSetAttribute(max dynamic shared memory): " + cu_error(r);
    r = a.FuncSetAttribute(c.fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, CU_SHAREDMEM_CARVEOUT_MAX_SHARED);
    if (r != CUDA_SUCCESS) return "cuFuncSetAttribute(carveout): " + cu_error(r);
    return "";
}

}  // namespace jit
}  // namespace bionc
