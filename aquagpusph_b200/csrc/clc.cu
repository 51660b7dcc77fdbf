// clc.cu -- run-time scripts: the `kernel` tool for a script that is NOT in the hand-written registry.
//
// The reference hands every script to the OpenCL compiler of the device at set-up
// (aquagpusph/CalcServer/Kernel.cpp:354-420: source + "-I<script folder> -I<base path>" + the problem's
// definitions, Tool.cpp:320-345: -DHAVE_2D/3D, usize = uint, -DNDEBUG; CalcServer.cpp:240-265: the -D list)
// and binds its arguments by name through clGetKernelArgInfo (Kernel.cpp:497-556).  The hot path of this
// library is hand-written (sweeps.cu, elementwise.cu, linklist.cu); what remains -- case-local scripts of the
// examples (init.cl, Rescale.cl, spring.cl, bc.cl, ...) and the boundary families nobody has written by hand
// yet -- takes THIS path: the script and the headers it includes are read where they lie,
// compiled for sm_100a by NVRTC behind a dialect header (clc_prelude.cuh), loaded with cudaLibraryLoadData and
// entered in the registry under the script's path, with the argument list parsed from its signature.  Its
// neighbour loops (BEGIN_NEIGHS of types/{2D,3D}.h) run as the reference wrote them, one particle per thread:
// correct, and as slow as a direct port -- a script that matters belongs in the registry.
// A registry kernel is never replaced by this path; libnvrtc is loaded on first use and its absence is an
// error at set-up, not a fallback to anything.
#include <dirent.h>
#include <dlfcn.h>

#include <algorithm>
#include <deque>
#include <fstream>
#include <regex>
#include <sstream>

#include "aqc_common.cuh"

namespace {

const char* const PRELUDE =
#include "clc_prelude.cuh"
    ;

// ---- libnvrtc, loaded on first use ---------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* so = nullptr;
    int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*);
    int (*CompileProgram)(nvrtcProgram, int, const char* const*);
    int (*GetProgramLogSize)(nvrtcProgram, size_t*);
    int (*GetProgramLog)(nvrtcProgram, char*);
    int (*GetCUBINSize)(nvrtcProgram, size_t*);
    int (*GetCUBIN)(nvrtcProgram, char*);
    int (*DestroyProgram)(nvrtcProgram*);
    const char* (*GetErrorString)(int);
};
Nvrtc* nvrtc(std::string& why)
{
    static Nvrtc n;
    static bool tried = false;
    if (!tried) {
        tried = true;
        for (const char* name : { "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12" }) {
            n.so = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (n.so)
                break;
        }
        if (n.so) {
#define SYM(f) *(void**)(&n.f) = dlsym(n.so, "nvrtc" #f)
            SYM(CreateProgram); SYM(CompileProgram); SYM(GetProgramLogSize); SYM(GetProgramLog);
            SYM(GetCUBINSize); SYM(GetCUBIN); SYM(DestroyProgram); SYM(GetErrorString);
#undef SYM
            if (!n.CreateProgram || !n.CompileProgram || !n.GetCUBIN || !n.GetCUBINSize) {
                dlclose(n.so);
                n.so = nullptr;
            }
        }
    }
    if (!n.so) {
        why = "libnvrtc.so.12 cannot be loaded: scripts outside the kernel registry need the CUDA run-time compiler";
        return nullptr;
    }
    return &n;
}

bool read_file(const std::string& path, std::string& out)
{
    std::ifstream f(path);
    if (!f)
        return false;
    std::ostringstream s;
    s << f.rdbuf();
    out = s.str();
    return true;
}

std::string dir_of(const std::string& p)
{
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? std::string(".") : p.substr(0, k);
}

// #include "file" expanded in place (the OpenCL compiler reads the headers from -I<script folder> and
// -I<base path>, Kernel.cpp:377-381): NVRTC could read them itself, but every header needs the literal rewrite
bool expand_includes(const std::string& text, const std::string& here, const std::vector<std::string>& roots,
                     int depth, std::string& out, std::string& err)
{
    if (depth > 32) {
        err = "#include nesting deeper than 32";
        return false;
    }
    static const std::regex inc(R"(^[ \t]*#[ \t]*include[ \t]*"([^"]+)\"[^\n]*$)");
    std::istringstream in(text);
    std::string line;
    while (std::getline(in, line)) {
        std::smatch m;
        if (std::regex_match(line, m, inc)) {
            std::vector<std::string> dirs{ here };
            dirs.insert(dirs.end(), roots.begin(), roots.end());
            std::string body, found;
            for (auto& d : dirs)
                if (read_file(d + "/" + m[1].str(), body)) {
                    found = d + "/" + m[1].str();
                    break;
                }
            if (found.empty()) {
                err = "cannot find the included file \"" + m[1].str() + "\"";
                return false;
            }
            out += "// >>> " + m[1].str() + "\n";
            if (!expand_includes(body, dir_of(found), roots, depth + 1, out, err))
                return false;
            out += "\n// <<< " + m[1].str() + "\n";
        } else {
            out += line;
            out += '\n';
        }
    }
    return true;
}

// "(float4)(a, b, c, d)" is a cast of a comma expression in C++: the OpenCL vector literal becomes a
// constructor call
std::string rewrite_literals(const std::string& s)
{
    static const std::regex lit(
        R"(\(\s*((?:float|double|int|uint|long|ulong|usize|ssize)(?:2|3|4|8|16)|(?:d|i|l|ui|ul|s|ss)?vec(?:2|3|4|8|_xyz)?|matrix)\s*\)\s*\()");
    return std::regex_replace(s, lit, "$1(");
}

std::string strip_comments(const std::string& s)
{
    std::string o;
    o.reserve(s.size());
    for (size_t k = 0; k < s.size();) {
        if (s.compare(k, 2, "/*") == 0) {
            const size_t e = s.find("*/", k + 2);
            k = e == std::string::npos ? s.size() : e + 2;
            o += ' ';
        } else if (s.compare(k, 2, "//") == 0) {
            const size_t e = s.find('\n', k);
            k = e == std::string::npos ? s.size() : e;
        } else {
            o += s[k++];
        }
    }
    return o;
}

struct ParsedArg {
    std::string name, type;
    int kind;
};

// the parameter list of `__kernel void entry(...)`: what clGetKernelArgInfo reported (Kernel.cpp:497-556)
bool parse_signature(const std::string& script, const std::string& entry, std::vector<ParsedArg>& out, std::string& err)
{
    const std::string s = strip_comments(script);
    const std::regex head("__kernel\\s+void\\s+" + entry + "\\s*\\(");
    std::smatch m;
    if (!std::regex_search(s, m, head)) {
        err = "no \"__kernel void " + entry + "(\" in the script";
        return false;
    }
    size_t k = (size_t)m.position(0) + (size_t)m.length(0);
    int depth = 1;
    std::string params;
    for (; k < s.size() && depth; k++) {
        if (s[k] == '(')
            depth++;
        else if (s[k] == ')')
            depth--;
        if (depth)
            params += s[k];
    }
    std::vector<std::string> items;
    std::string cur;
    depth = 0;
    for (char c : params) {
        if (c == '(')
            depth++;
        if (c == ')')
            depth--;
        if (c == ',' && !depth) {
            items.push_back(cur);
            cur.clear();
        } else {
            cur += c;
        }
    }
    items.push_back(cur);
    auto push = [&](const char* n, const char* t, int kind) { out.push_back({ n, t, kind }); };
    for (auto& raw : items) {
        std::istringstream ts(raw);
        std::vector<std::string> tok;
        std::string t;
        // split at blanks and around '*'
        std::string spaced;
        for (char c : raw) {
            if (c == '*')
                spaced += " * ";
            else
                spaced += c;
        }
        std::istringstream sp(spaced);
        while (sp >> t)
            tok.push_back(t);
        if (tok.empty())
            continue;
        // types.h:106-122
        if (tok.size() == 1 && tok[0] == "LINKLIST_LOCAL_PARAMS") {
            push("icell", "usize*", AQC_ARG_ARRAY_IN);
            push("ihoc", "usize*", AQC_ARG_ARRAY_IN);
            push("n_cells", "svec4", AQC_ARG_SCALAR);
            continue;
        }
        if (tok.size() == 1 && tok[0] == "LINKLIST_REMOTE_PARAMS") {
            push("icell", "usize*", AQC_ARG_ARRAY_IN);
            push("mpi_icell", "usize*", AQC_ARG_ARRAY_IN);
            push("mpi_ihoc", "usize*", AQC_ARG_ARRAY_IN);
            push("n_cells", "svec4", AQC_ARG_SCALAR);
            continue;
        }
        ParsedArg a;
        a.name = tok.back();
        tok.pop_back();
        bool pointer = false, is_const = false;
        std::string type;
        for (auto& w : tok) {
            if (w == "*")
                pointer = true;
            else if (w == "const")
                is_const = is_const || !pointer; // (const before the '*': the pointee)
            else if (w == "__global" || w == "__constant" || w == "restrict" || w == "__restrict" ||
                     w == "__private" || w == "__local")
                is_const = is_const || w == "__constant";
            else
                type += (type.empty() ? "" : " ") + w;
        }
        if (type == "uint")
            type = "unsigned int";
        if (type == "size_t")
            type = "usize";
        a.type = type + (pointer ? "*" : "");
        a.kind = !pointer ? AQC_ARG_SCALAR : (is_const ? AQC_ARG_ARRAY_IN : AQC_ARG_ARRAY_OUT);
        out.push_back(a);
    }
    return true;
}

struct Script {
    std::string key;    // path::entry::dims::defines
    cudaLibrary_t lib = nullptr;
    cudaKernel_t fn = nullptr;
    std::vector<size_t> scalar_bytes; // per argument; 0 for arrays
    int id = -1;
};
std::deque<Script>& scripts()
{
    static std::deque<Script> s;
    return s;
}
std::deque<std::string>& strings() // stable storage of the names the registry points at
{
    static std::deque<std::string> s;
    return s;
}
const char* keep(const std::string& s)
{
    strings().push_back(s);
    return strings().back().c_str();
}

size_t scalar_size(const std::string& type, int dims)
{
    if (type == "matrix")
        return dims == 3 ? 64 : 16;
    return aqc_type_bytes(type.c_str(), dims);
}

} // namespace

// registry.cu side: launch of a run-time script (aqc_launch)
int aqc_script_launch(aqc_ctx* ctx, const aqc_kernel_entry& e, size_t n, void* const* args)
{
    const Script* sc = (const Script*)e.jit;
    std::vector<void*> slots(e.args.size());
    std::vector<void*> ptrs(e.args.size());
    for (size_t k = 0; k < e.args.size(); k++) {
        if (e.args[k].kind == AQC_ARG_SCALAR) {
            slots[k] = args[k]; // host pointer to the value
        } else {
            ptrs[k] = args[k]; // the device pointer itself is the value
            slots[k] = &ptrs[k];
        }
    }
    const unsigned block = 256; // = LOCAL_MEM_SIZE of the compilation
    AQC_CUDA(ctx, cudaLaunchKernel((const void*)sc->fn, dim3(aqc_blocks(n, block)), dim3(block), slots.data(), 0,
                                   ctx->stream));
    AQC_LAUNCH_CHECK(ctx);
    return AQC_OK;
}

// script -> cubin + argument list (no device needed)
static int build_cubin(const char* path, const std::string& ent, int dims, const char* base_path,
                       const char* const* defines, int ndefines, std::vector<char>& cubin,
                       std::vector<ParsedArg>& pargs, std::string& err)
{
    std::string src;
    if (!read_file(path, src)) {
        err = "cannot read the script";
        return AQC_ERR_NOKERNEL;
    }
    Nvrtc* N = nvrtc(err);
    if (!N)
        return AQC_ERR_NOKERNEL;
    std::vector<std::string> roots;
    if (base_path && *base_path)
        roots.push_back(base_path);
    std::string body;
    if (!expand_includes(src, dir_of(path), roots, 0, body, err))
        return AQC_ERR_NOKERNEL;
    body = rewrite_literals(body);
    if (!parse_signature(body, ent, pargs, err))
        return AQC_ERR_NOKERNEL;
    // Tool.cpp:328-345 + Kernel.cpp:411: the flags every script is compiled with
    std::string full = std::string(PRELUDE) + "\n#define NDEBUG\n#define X32\n#define LOCAL_MEM_SIZE 256\n" +
                       (dims == 3 ? "#define HAVE_3D\n" : "#define HAVE_2D\n");
    for (int k = 0; k < ndefines; k++) { // "-DNAME=VALUE" / "-DNAME" (CalcServer.cpp:240-265)
        std::string d = defines[k];
        if (d.compare(0, 2, "-D") == 0)
            d = d.substr(2);
        const size_t eq = d.find('=');
        full += "#define " + (eq == std::string::npos ? d : d.substr(0, eq) + " " + d.substr(eq + 1)) + "\n";
    }
    full += "#line 1 \"" + std::string(path) + "\"\n" + body;
    // KernelFunctions/Kernel.h includes the kernel's file through a macro (#include KERNEL_STRINGIFY(...
    // KERNEL_NAME ...hcl)), which only the compiler's own preprocessor can resolve: the few files of that
    // folder travel as in-memory headers, rewritten like everything else
    std::vector<std::string> hnames, hbodies;
    for (auto& root : roots) {
        const std::string dir = root + "/resources/Scripts/KernelFunctions";
        std::vector<std::string> files;
        if (DIR* dh = opendir(dir.c_str())) {
            while (dirent* de = readdir(dh)) {
                const std::string fn = de->d_name;
                if (fn.size() > 4 && fn.compare(fn.size() - 4, 4, ".hcl") == 0)
                    files.push_back(fn);
            }
            closedir(dh);
        }
        std::sort(files.begin(), files.end());
        for (auto& fn : files) {
            std::string raw, exp;
            if (!read_file(dir + "/" + fn, raw))
                continue;
            if (!expand_includes(raw, dir, roots, 1, exp, err))
                return AQC_ERR_NOKERNEL;
            hnames.push_back(std::string("resources/Scripts/KernelFunctions/") + fn);
            hbodies.push_back(rewrite_literals(exp));
        }
    }
    std::vector<const char*> hn, hb;
    for (size_t k = 0; k < hnames.size(); k++) {
        hn.push_back(hnames[k].c_str());
        hb.push_back(hbodies[k].c_str());
    }
    if (const char* dump = getenv("AQC_SCRIPT_DUMP_SRC")) { // diagnostics: what the compiler is given
        std::ofstream f(dump);
        f << full;
    }
    nvrtcProgram prog = nullptr;
    int rc = N->CreateProgram(&prog, full.c_str(), "script.cu", (int)hn.size(), hb.data(), hn.data());
    if (rc) {
        err = std::string("nvrtcCreateProgram: ") + (N->GetErrorString ? N->GetErrorString(rc) : "?");
        return AQC_ERR_CUDA;
    }
    // (no FMA contraction: element-wise scripts then give the same bits as on any IEEE device)
    const char* opts[] = { "--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-default-device", "-w" };
    rc = N->CompileProgram(prog, 5, opts);
    if (rc) {
        size_t ls = 0;
        N->GetProgramLogSize(prog, &ls);
        std::string log(ls ? ls : 1, '\0');
        if (ls)
            N->GetProgramLog(prog, &log[0]);
        N->DestroyProgram(&prog);
        // the first error lines are what a user needs
        const size_t e = log.find("error");
        if (e != std::string::npos && e > 200)
            log = log.substr(e - 200);
        if (log.size() > 420)
            log.resize(420);
        err = "does not compile: " + log;
        return AQC_ERR_NOKERNEL;
    }
    size_t cs = 0;
    N->GetCUBINSize(prog, &cs);
    cubin.resize(cs);
    N->GetCUBIN(prog, cubin.data());
    N->DestroyProgram(&prog);
    if (const char* dump = getenv("AQC_SCRIPT_DUMP")) { // diagnostics: the cubin of the last script
        std::ofstream f(dump, std::ios::binary);
        f.write(cubin.data(), (std::streamsize)cubin.size());
    }
    return AQC_OK;
}

// compile only (no device, nothing registered): the number of arguments, or < 0 with the reason in `log`
extern "C" int aqc_script_check(const char* path, const char* entry, int dims, const char* base_path,
                                const char* const* defines, int ndefines, char* log, size_t log_bytes)
{
    if (!path)
        return AQC_ERR_ARG;
    std::vector<char> cubin;
    std::vector<ParsedArg> pargs;
    std::string err;
    const int rc = build_cubin(path, (entry && *entry) ? entry : "entry", dims, base_path, defines, ndefines, cubin,
                               pargs, err);
    if (log && log_bytes) {
        std::string names;
        for (auto& a : pargs)
            names += a.type + " " + a.name + (a.kind == AQC_ARG_ARRAY_OUT ? " (out); " : "; ");
        snprintf(log, log_bytes, "%s", rc ? err.c_str() : names.c_str());
    }
    return rc ? rc : (int)pargs.size();
}

extern "C" int aqc_script_compile(aqc_ctx* ctx, const char* path, const char* entry, int dims, const char* base_path,
                                  const char* const* defines, int ndefines)
{
    if (!ctx || !path)
        return AQC_ERR_ARG;
    const std::string ent = (entry && *entry) ? entry : "entry";
    std::string key = std::string(path) + "::" + ent + "::" + std::to_string(dims);
    for (int k = 0; k < ndefines; k++)
        key += std::string(" ") + defines[k];
    for (auto& s : scripts())
        if (s.key == key)
            return s.id;
    std::vector<char> cubin;
    std::vector<ParsedArg> pargs;
    std::string err;
    if (int rc = build_cubin(path, ent, dims, base_path, defines, ndefines, cubin, pargs, err))
        return aqc_fail(ctx, rc, "script \"%s\": %s", path, err.c_str());
    Script sc;
    sc.key = key;
    AQC_CUDA(ctx, cudaLibraryLoadData(&sc.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
    AQC_CUDA(ctx, cudaLibraryGetKernel(&sc.fn, sc.lib, ent.c_str()));
    aqc_kernel_entry e;
    // the registry's key: what follows the last "Scripts/" of the path, else the whole path
    std::string rel = path;
    for (size_t p = rel.find("Scripts/"); p != std::string::npos; p = rel.find("Scripts/"))
        rel = rel.substr(p + 8);
    e.script = keep(rel);
    e.entry = keep(ent);
    e.dims = dims;
    for (auto& a : pargs) {
        aqc_arg_info ai;
        ai.name = keep(a.name);
        ai.type = keep(a.type);
        ai.kind = a.kind;
        e.args.push_back(ai);
        sc.scalar_bytes.push_back(a.kind == AQC_ARG_SCALAR ? scalar_size(a.type, dims) : 0);
    }
    e.fn = nullptr;
    sc.id = (int)aqc_registry().size();
    scripts().push_back(sc);
    e.jit = &scripts().back();
    aqc_registry().push_back(e);
    return sc.id;
}
