// savers.cpp -- particle output files and the restart checkpoint of a run.
//
// Counterpart of aquagpusph/InputOutput/Particles.cpp:82-120, 243-323 (save + download on a
// parallel command queue, file written from an event callback), ASCII.cpp:240-333 (the .dat layout),
// AuxiliarMethods.cpp:215-256 (file numbering), FileManager.cpp:146-164 and
// State.cpp:195-226, 1517-1908 (the AQUAgpusph.save.N.xml state file a run can be resumed from).
//
// B200 design: every <Save> field is un-sorted ON THE DEVICE into a staging buffer of its own (one
// scatter launch for all fields of a set on the main stream), copied to pinned host memory on the
// context's side stream (aqc_side_*: the copies overlap the next time steps), and a writer thread
// waits for the copy event and formats the file -- the main stream never waits for a file.
#include <atomic>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <limits>
#include <thread>

#include "calcserver.hpp"

namespace fs = std::filesystem;

namespace Aqua {
namespace CalcServer {

using InputOutput::ProblemSetup;
using InputOutput::Variable;

struct ParticlesSaver::Field {
    Variable* var = nullptr;
    void* dev = nullptr;  // un-sorted copy (whole array)
    void* host = nullptr; // pinned, the set's rows only
    size_t bytes = 0;     // of the set's rows
};

ParticlesSaver::ParticlesSaver(CalcServer* C, size_t iset, size_t first, size_t n,
                               const std::string& path, const std::string& format,
                               const std::string& fields)
  : _C(C), _iset(iset), _first(first), _n(n), _path(path), _fields_txt(fields)
{
    const std::string f = toLowerCopy(format);
    if (f == "ascii" || f == "fastascii") {
        _format = format;
        _ext = ".dat";
    } else if (f == "csv") {
        _format = format;
        _ext = ".csv";
        _sep = ',';
        _csep = ',';
    } else {
        // VTK needs libvtk in the reference (VTK.cpp) and is outside this host: the same fields go
        // to a FastASCII file, which the checkpoint then names as the set's <Load>
        log(L_WARNING, "Particles set " + std::to_string(iset) + ": output format \"" + format +
                           "\" is not provided by this host, writing FastASCII instead\n");
        _format = "FastASCII";
        _ext = ".dat";
    }
    for (auto name : split(fields)) {
        if (name.empty())
            continue;
        Variable* v = C->variables()->get(name);
        if (!v)
            throw std::runtime_error("Can't download undeclared variable \"" + name + "\".");
        if (!v->isArray())
            throw std::runtime_error("Variable \"" + name + "\" is a scalar.");
        if (v->length() < first + n)
            throw std::runtime_error("Variable \"" + name + "\" is not long enough.");
        auto fld = std::make_unique<Field>();
        fld->var = v;
        _fields.push_back(std::move(fld));
    }
    if (_fields.empty())
        throw std::runtime_error("No fields have been marked to be saved");
    if (aqc_event_create(C->ctx(), &_event))
        throw std::runtime_error(aqc_last_error(C->ctx()));
}

ParticlesSaver::~ParticlesSaver()
{
    try {
        wait();
    } catch (...) {
    }
    for (auto& f : _fields) {
        if (f->dev)
            aqc_free(_C->ctx(), f->dev);
        if (f->host)
            aqc_host_free(_C->ctx(), f->host);
    }
    if (_event)
        aqc_event_destroy(_C->ctx(), _event);
}

void ParticlesSaver::wait()
{
    if (_writer.joinable())
        _writer.join();
    if (!_error.empty()) {
        const std::string e = _error;
        _error.clear();
        throw std::runtime_error(e);
    }
}

// AuxiliarMethods.cpp:215-256: "%d" is the old spelling of "{index}"; without a place for the index
// ".{index}<ext>" is appended (ASCII.cpp:462-468); the first free index from the last one used
std::string ParticlesSaver::nextFile()
{
    std::string base = replaceAllCopy(_path, "%d", "{index}");
    if (base.find("{index}") == std::string::npos)
        base += ".{index}" + _ext;
    while (true) {
        const std::string p = formatPath(base, _C->mpi_rank(), (int)_next_index);
        std::error_code ec;
        if (!fs::exists(p, ec))
            return p;
        _next_index++;
    }
}

void ParticlesSaver::save(float t)
{
    wait(); // "Just one instance at a time" (Particles.cpp:93-94)
    aqc_ctx* ctx = _C->ctx();
    auto chk = [&](int rc) {
        if (rc)
            throw std::runtime_error(aqc_last_error(ctx));
    };
    Variable* id = _C->variables()->get("id");
    // un-sort: out[id[i]] = in[i] (UnSort.cl.in:30-42), every per-particle field in one launch
    std::vector<const void*> src;
    std::vector<void*> dst;
    std::vector<size_t> eb;
    for (auto& f : _fields) {
        Variable* v = f->var;
        const size_t set_bytes = _n * v->typesize();
        if (f->bytes != set_bytes) {
            if (f->host)
                chk(aqc_host_free(ctx, f->host));
            if (f->dev)
                chk(aqc_free(ctx, f->dev));
            f->host = f->dev = nullptr;
            chk(aqc_host_alloc(ctx, set_bytes ? set_bytes : 1, &f->host));
            f->bytes = set_bytes;
        }
        if (v->length() == id->length()) {
            if (!f->dev)
                chk(aqc_alloc(ctx, v->size(), &f->dev));
            src.push_back(v->dptr());
            dst.push_back(f->dev);
            eb.push_back(v->typesize());
        }
    }
    if (!src.empty())
        chk(aqc_scatter_fields(ctx, (const aqc_usize*)id->dptr(), (aqc_usize)id->length(), (int)src.size(),
                               src.data(), dst.data(), eb.data()));
    // the side stream takes over: it waits for the scatter, copies, and signals the writer
    chk(aqc_side_fork(ctx));
    for (auto& f : _fields) {
        Variable* v = f->var;
        const char* from = (const char*)(f->dev ? f->dev : v->dptr()) + _first * v->typesize();
        chk(aqc_memcpy_d2h_side(ctx, f->host, from, f->bytes));
    }
    chk(aqc_side_record(ctx, _event));
    _file = nextFile();
    _next_index++;
    // reserve the name now: the next set / the checkpoint may ask for file() before the writer ran
    { std::ofstream touch(_file); }
    _time = t;
    _writer = std::thread([this]() {
        try {
            if (aqc_side_wait(_C->ctx(), _event))
                throw std::runtime_error("the download of particles set " + std::to_string(_iset) + " failed");
            print_file();
        } catch (std::exception& e) {
            _error = e.what();
        }
    });
}

static void print_component(std::string& out, char kind, const char* p)
{
    char b[64];
    switch (kind) {
        case 'i': snprintf(b, sizeof(b), "%d", *(const int32_t*)p); break;
        case 'u': snprintf(b, sizeof(b), "%u", *(const uint32_t*)p); break;
        case 'l': snprintf(b, sizeof(b), "%lld", (long long)*(const int64_t*)p); break;
        case 'L': snprintf(b, sizeof(b), "%llu", (unsigned long long)*(const uint64_t*)p); break;
        case 'd': snprintf(b, sizeof(b), "%.17g", *(const double*)p); break;
        // ASCII.cpp:252-260 sets digits10 + 1 = 7 significant digits, which does not round-trip a
        // float; 9 (max_digits10) does, so that a run resumed from the file continues from the very
        // same state.  Every reader of the format accepts both.
        default: snprintf(b, sizeof(b), "%.9g", (double)*(const float*)p);
    }
    out += b;
}

void ParticlesSaver::print_file()
{
    FILE* f = fopen(_file.c_str(), "w");
    if (!f)
        throw std::runtime_error("Cannot write \"" + _file + "\"");
    // ASCII.cpp:311-333
    fprintf(f, "#########################################################\n"
               "#\n"
               "#    File autogenerated by AQUAgpusph-b200\n"
               "#    t = %.9g s\n"
               "#    fields = %s\n"
               "#\n"
               "#########################################################\n\n",
            (double)_time, _fields_txt.c_str());
    std::string line;
    for (size_t i = 0; i < _n; i++) {
        line.clear();
        for (size_t k = 0; k < _fields.size(); k++) {
            const Variable* v = _fields[k]->var;
            const char* e = (const char*)_fields[k]->host + i * v->typesize();
            const unsigned nc = v->ncomp();
            for (unsigned c = 0; c < nc; c++) {
                print_component(line, v->kind(), e + c * v->compsize());
                if (c + 1 < nc)
                    line += _csep;
            }
            if (k + 1 < _fields.size())
                line += _sep;
        }
        line += '\n';
        fputs(line.c_str(), f);
    }
    if (fclose(f))
        throw std::runtime_error("Failure writing \"" + _file + "\"");
    log(L_INFO, "Wrote \"" + _file + "\" ASCII file.\n");
}

// ------------------------------------------------------------------ CalcServer --
void CalcServer::save(float t)
{
    if (_savers.empty()) {
        size_t first = 0;
        for (size_t i = 0; i < _sim_data.sets.size(); i++) {
            auto& set = *_sim_data.sets[i];
            for (auto& o : set.outputs)
                _savers.emplace_back(new ParticlesSaver(this, i, first, set.n, o[0], o[1], o[2]));
            first += set.n;
        }
    }
    // FileManager::save (FileManager.cpp:146-155): the savers, then the XML definition file
    for (auto& s : _savers)
        s->save(t);
    writeCheckpoint();
}

void CalcServer::waitForSavers()
{
    for (auto& s : _savers)
        s->wait();
}

// State.cpp:195-226 (first free AQUAgpusph.save.N.xml) and :1517-1908 (what it holds): the problem as
// it stands NOW -- every variable with its current value / length, the tools, the timing options and,
// per particles set, a <Load> of the file just written next to the unchanged <Save>.
void CalcServer::writeCheckpoint()
{
    if (_checkpoint_file.empty()) {
        unsigned i = 0;
        std::error_code ec;
        std::string name;
        do {
            name = "AQUAgpusph.save." + std::to_string(i++) + ".xml";
            if (_mpi_size > 1)
                name = "AQUAgpusph.save." + std::to_string(i - 1) + ".rank" + std::to_string(_mpi_rank) + ".xml";
        } while (fs::exists(name, ec));
        _checkpoint_file = name;
    }
    ProblemSetup sd = _sim_data; // (tools / reports are shared pointers: not modified here)
    sd.variables.clear();
    for (auto& v : _vars->all()) {
        if (startswith(v->name(), "__"))
            continue;
        if (v->isArray()) {
            sd.registerVariable(v->name(), v->type(), std::to_string(v->length()), "");
        } else {
            std::string txt = v->asString();
            if (!txt.empty() && txt.front() == '(')
                txt.front() = ' ';
            if (!txt.empty() && txt.back() == ')')
                txt.back() = ' ';
            sd.registerVariable(v->name(), v->type(), "", trimCopy(txt));
        }
    }
    sd.sets.clear();
    for (size_t i = 0; i < _sim_data.sets.size(); i++) {
        auto set = std::make_shared<ProblemSetup::ParticlesSet>(*_sim_data.sets[i]);
        set->n_known = true;
        for (auto& kv : set->scalars) { // the set's CURRENT value of every per-set scalar
            Variable* v = _vars->get(kv.first);
            if (!v || !v->isArray() || v->length() <= i)
                continue;
            std::vector<char> data(v->typesize());
            if (aqc_memcpy_d2h(_ctx, data.data(), (const char*)v->dptr() + i * v->typesize(), v->typesize(), 1))
                throw std::runtime_error(aqc_last_error(_ctx));
            std::string txt;
            for (unsigned c = 0; c < v->ncomp(); c++) {
                if (c)
                    txt += ",";
                print_component(txt, v->kind(), data.data() + c * v->compsize());
            }
            kv.second = txt;
        }
        for (auto& s : _savers)
            if (s->set() == i) { // (the first <Save> of the set is what a resumed run loads)
                set->in_path = fs::absolute(s->file()).string();
                set->in_format = s->format();
                set->in_fields = s->fields();
                break;
            }
        sd.sets.push_back(set);
    }
    InputOutput::State().write(_checkpoint_file, sd, false, true);
    log(L_INFO, "Wrote \"" + _checkpoint_file + "\" SPH state file...\n");
}

} // namespace CalcServer
} // namespace Aqua
