// problem.hpp -- in-memory description of a simulation (what the XML files say)
// and the XML front-end that fills it.  Counterpart of
// aquagpusph/InputOutput/{ProblemSetup,State}.*: same tags, attributes,
// defaults and tool placement rules (State.cpp:276-389, 739-1217), so the
// reference's case files and preset pipelines are accepted unchanged.
#pragma once
#include <array>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "aux.hpp"
#include "xml.hpp"

namespace Aqua {
namespace InputOutput {

struct ProblemSetup {
    struct Settings {
        bool save_on_fail = true;
        std::string base_path;
        struct Device { unsigned platform = 0, device = 0; std::string type = "ALL"; unsigned addr_bits = 32; };
        std::vector<Device> devices;
    } settings;

    struct VariableDef { std::string name, type, length, value; };
    std::vector<VariableDef> variables;
    // duplicates are kept, in document order: CalcServer registers them one by one
    // and a later registration replaces the earlier variable (ProblemSetup.cpp:100-109,
    // Variable.cpp:1058-1064)
    void registerVariable(const std::string& n, const std::string& t, const std::string& l,
                          const std::string& v)
    {
        variables.push_back({ n, t, l, v });
    }

    struct Definition { std::string name, value; bool evaluate; };
    std::vector<Definition> definitions;
    void define(const std::string& n, const std::string& v, bool e)
    {
        // re-defining moves the definition to the end (ProblemSetup.cpp:112-123)
        for (size_t i = 0; i < definitions.size(); i++)
            if (definitions[i].name == n) {
                definitions.erase(definitions.begin() + i);
                break;
            }
        definitions.push_back({ n, v, e });
    }
    bool isDefined(const std::string& n) const
    {
        for (auto& d : definitions)
            if (d.name == n)
                return true;
        return false;
    }

    /// A tool or report: a bag of string attributes (ProblemSetup::sphTool)
    struct Tool {
        std::vector<std::pair<std::string, std::string>> data;
        void set(const std::string& k, const std::string& v)
        {
            for (auto& kv : data)
                if (kv.first == k) {
                    kv.second = v;
                    return;
                }
            data.emplace_back(k, v);
        }
        bool has(const std::string& k) const
        {
            for (auto& kv : data)
                if (kv.first == k)
                    return true;
            return false;
        }
        std::string get(const std::string& k) const
        {
            for (auto& kv : data)
                if (kv.first == k)
                    return kv.second;
            return "";
        }
    };
    // the same tool may sit at several places of the pipeline (wildcard inserts)
    std::vector<std::shared_ptr<Tool>> tools;
    std::vector<std::shared_ptr<Tool>> reports;

    struct TimeOpts {
        enum { NO_OUTPUT = 0, TIME_MODE = 1, ITER_MODE = 2, FRAME_MODE = 4, FPS_MODE = 1, IPF_MODE = 2 };
        unsigned sim_end_mode = 0;
        float sim_end_time = 0.f;
        unsigned sim_end_step = 0, sim_end_frame = 0;
        unsigned output_mode = 0;
        float output_fps = 0.f;
        unsigned output_ipf = 0;
    } time_opts;

    struct ParticlesSet {
        size_t n = 0;
        bool n_known = false;
        std::vector<std::pair<std::string, std::string>> scalars;
        std::string in_path, in_format, in_fields;
        std::vector<std::array<std::string, 3>> outputs; // path, format, fields
    };
    std::vector<std::shared_ptr<ParticlesSet>> sets;

    int dims = 3;
};

class State {
  public:
    /// Parse `input_file` (and everything it includes) into sim_data
    void load(const std::string& input_file, ProblemSetup& sim_data);
    /// Write the fully resolved problem (one flat XML, no includes): the format of
    /// the reference's AQUAgpusph.save.NNNNN.xml checkpoints (State.cpp:1517-1908)
    /// full_load_paths: <Load file=...> keeps its directory (a checkpoint names the file a saver
    /// just wrote); otherwise only the file name, as the examples' templates have it
    void write(const std::string& output_file, const ProblemSetup& sim_data,
               bool relative_script_paths = true, bool full_load_paths = false) const;

  private:
    void parse(const std::string& filepath, ProblemSetup& sim_data, const std::string& prefix);
    std::string findPath(const std::string& filepath, const ProblemSetup& sim_data,
                         bool must_exist) const;
    void parseSettings(const Xml::Node* root, ProblemSetup& sim_data);
    void parseVariables(const Xml::Node* root, ProblemSetup& sim_data);
    void parseDefinitions(const Xml::Node* root, ProblemSetup& sim_data);
    void parseTools(const Xml::Node* root, ProblemSetup& sim_data, const std::string& prefix);
    void parseReports(const Xml::Node* root, ProblemSetup& sim_data, const std::string& prefix);
    void parseTiming(const Xml::Node* root, ProblemSetup& sim_data);
    void parseSets(const Xml::Node* root, ProblemSetup& sim_data);
    void configureTool(ProblemSetup::Tool* tool, const Xml::Node* e, ProblemSetup& sim_data);
    std::vector<std::string> _xml_paths;
};

} // namespace InputOutput
} // namespace Aqua
