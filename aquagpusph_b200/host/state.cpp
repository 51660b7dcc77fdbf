// state.cpp -- XML front-end (see problem.hpp).  Follows the parsing order and
// tool-placement rules of aquagpusph/InputOutput/State.cpp:276-389, 739-1217.
#include <filesystem>
#include <fstream>

#include "problem.hpp"

namespace fs = std::filesystem;

namespace Aqua {
namespace InputOutput {

void State::load(const std::string& input_file, ProblemSetup& sim_data)
{
    parse(input_file, sim_data, "");
}

// State.cpp:234-274: as given / cwd, then <RootPath>, then the folder of the XML
// file being parsed.  Script paths need not exist here (kernels are pre-built
// CUDA, looked up by the path after ".../Scripts/"), so they are normalised
// lexically when the file is absent.
std::string State::findPath(const std::string& filepath, const ProblemSetup& sim_data,
                            bool must_exist) const
{
    const fs::path fp(filepath);
    std::error_code ec;
    if (fs::exists(fp, ec))
        return fs::canonical(fp).string();
    std::vector<std::string> candidates = { (fs::current_path() / fp).string() };
    if (fp.is_relative()) {
        if (!sim_data.settings.base_path.empty()) {
            const fs::path f = fs::path(sim_data.settings.base_path) / fp;
            if (fs::exists(f, ec))
                return fs::canonical(f).string();
            candidates.push_back(f.string());
        }
        if (!_xml_paths.empty()) {
            const fs::path f = fs::path(_xml_paths.back()) / fp;
            if (fs::exists(f, ec))
                return fs::canonical(f).string();
            candidates.push_back(f.string());
            if (!must_exist)
                return f.lexically_normal().string();
        }
    }
    if (!must_exist)
        return fp.lexically_normal().string();
    std::string msg = "No such file or directory '" + filepath + "'. Checked:";
    for (auto& c : candidates)
        msg += "\n  '" + c + "'";
    throw std::runtime_error(msg);
}

void State::parse(const std::string& filepath, ProblemSetup& sim_data, const std::string& prefix)
{
    log(L_INFO, "Parsing the XML file \"" + filepath + "\" with prefix \"" + prefix + "\"");
    auto doc = Xml::parseFile(filepath);
    const Xml::Node* root = Xml::root(doc.get());
    if (!root)
        throw std::runtime_error("Empty XML file " + filepath);
    _xml_paths.push_back(fs::path(filepath).parent_path().string());

    auto includes = [&](bool at_end) {
        for (const Xml::Node* e : root->descendants("Include")) {
            const bool when_end = e->has("when") && e->attr("when") == "end";
            const bool when_begin = !e->has("when") || e->attr("when") == "begin";
            if (at_end ? !when_end : !when_begin)
                continue;
            const std::string file = findPath(trimCopy(e->attr("file")), sim_data, true);
            const std::string pre = e->has("prefix") ? e->attr("prefix") : prefix;
            parse(file, sim_data, pre);
        }
    };
    includes(false);
    parseSettings(root, sim_data);
    parseVariables(root, sim_data);
    parseDefinitions(root, sim_data);
    parseTools(root, sim_data, prefix);
    parseReports(root, sim_data, prefix);
    parseTiming(root, sim_data);
    parseSets(root, sim_data);
    includes(true);
    _xml_paths.pop_back();
}

void State::parseSettings(const Xml::Node* root, ProblemSetup& sim_data)
{
    for (const Xml::Node* e : root->descendants("Settings")) {
        for (const Xml::Node* s : e->descendants("SaveOnFail"))
            sim_data.settings.save_on_fail = toLowerCopy(s->attr("value")) == "true";
        for (const Xml::Node* s : e->descendants("RootPath")) {
            if (!s->has("path"))
                throw std::runtime_error("RootPath without \"path\" attribute");
            try {
                sim_data.settings.base_path = findPath(s->attr("path"), sim_data, true);
            } catch (std::exception& ex) {
                log(L_WARNING, std::string("Ignoring RootPath: ") + ex.what());
            }
        }
        for (const Xml::Node* s : e->descendants("Device")) {
            ProblemSetup::Settings::Device d;
            d.platform = s->has("platform") ? std::stoi(s->attr("platform")) : 0;
            d.device = s->has("device") ? std::stoi(s->attr("device")) : 0;
            if (s->has("type"))
                d.type = s->attr("type");
            if (s->has("addr_bits"))
                d.addr_bits = std::stoi(s->attr("addr_bits"));
            sim_data.settings.devices.push_back(d);
        }
    }
}

void State::parseVariables(const Xml::Node* root, ProblemSetup& sim_data)
{
    for (const Xml::Node* e : root->descendants("Variables"))
        for (const Xml::Node* s : e->descendants("Variable")) {
            if (!s->has("name"))
                throw std::runtime_error("Found a variable without name");
            const std::string name = s->attr("name");
            if (startswith(name, "__"))
                throw std::runtime_error("Invalid variable name \"" + name +
                                         "\": prefix \"__\" is reserved");
            for (auto suf : { "_x", "_y", "_z", "_w" })
                if (endswith(name, suf))
                    throw std::runtime_error("Invalid variable name \"" + name + "\": suffix \"" +
                                             suf + "\" is reserved");
            if (s->attr("type").find('*') == std::string::npos)
                sim_data.registerVariable(name, s->attr("type"), "1", s->attr("value"));
            else
                sim_data.registerVariable(name, s->attr("type"), s->attr("length"), "");
        }
}

void State::parseDefinitions(const Xml::Node* root, ProblemSetup& sim_data)
{
    for (const Xml::Node* e : root->descendants("Definitions"))
        for (const Xml::Node* s : e->descendants("Define")) {
            if (!s->has("name"))
                throw std::runtime_error("Name shall be specified for definitions");
            if (!s->has("value")) {
                sim_data.define(s->attr("name"), "", false);
                continue;
            }
            const std::string ev = toLowerCopy(s->attr("evaluate"));
            sim_data.define(s->attr("name"), s->attr("value"), ev == "true" || ev == "yes");
        }
}

static std::vector<unsigned> toolsList(const std::string& list, const ProblemSetup& sd,
                                       const std::string& prefix)
{
    std::vector<unsigned> places;
    std::istringstream f(list);
    std::string s;
    while (std::getline(f, s, ',')) // names are NOT trimmed (State.cpp:650-668)
        for (unsigned p = 0; p < sd.tools.size(); p++)
            if (prefix + s == sd.tools[p]->get("name"))
                places.push_back(p);
    return places;
}

static std::vector<unsigned> toolsName(const std::string& name, const ProblemSetup& sd,
                                       const std::string& prefix)
{
    std::vector<unsigned> places;
    for (unsigned p = 0; p < sd.tools.size(); p++)
        if (match(prefix + name, sd.tools[p]->get("name")))
            places.push_back(p);
    return places;
}

static void toolAttr(ProblemSetup::Tool* t, const Xml::Node* e, const std::string& a,
                     const std::string& def)
{
    t->set(a, e->has(a) ? e->attr(a) : def);
}

static void toolAttr(ProblemSetup::Tool* t, const Xml::Node* e, const std::string& a)
{
    if (!e->has(a))
        throw std::runtime_error("Tool \"" + t->get("name") + "\" requires attribute \"" + a + "\"");
    t->set(a, e->attr(a));
}

void State::configureTool(ProblemSetup::Tool* tool, const Xml::Node* e, ProblemSetup& sd)
{
    const std::string type = e->attr("type");
    if (type == "kernel") {
        toolAttr(tool, e, "path");
        tool->set("path", findPath(e->attr("path"), sd, false));
        toolAttr(tool, e, "entry_point", "entry");
        toolAttr(tool, e, "n", "");
    } else if (type == "copy") {
        toolAttr(tool, e, "in");
        toolAttr(tool, e, "out");
    } else if (type == "python" || type == "installable") {
        toolAttr(tool, e, "path");
        tool->set("path", findPath(e->attr("path"), sd, false));
    } else if (type == "set" || type == "set_scalar") {
        toolAttr(tool, e, "in");
        toolAttr(tool, e, "value");
    } else if (type == "reduction") {
        for (auto a : { "in", "out", "null" })
            toolAttr(tool, e, a);
        if (trimCopy(e->text).empty())
            throw std::runtime_error("No operation specified for the reduction \"" +
                                     tool->get("name") + "\"");
        tool->set("operation", e->text);
    } else if (type == "link-list") {
        toolAttr(tool, e, "in", "r");
        toolAttr(tool, e, "min", "r_min");
        toolAttr(tool, e, "max", "r_max");
        toolAttr(tool, e, "ihoc", "ihoc");
        toolAttr(tool, e, "icell", "icell");
        toolAttr(tool, e, "n_cells", "n_cells");
        toolAttr(tool, e, "perm", "id_unsorted");
        toolAttr(tool, e, "inv_perm", "id_sorted");
        toolAttr(tool, e, "recompute_grid", "true");
        toolAttr(tool, e, "sorter", "radix-sort");
        // (ours, optional) arrays whose writes are the only way the input positions can change:
        // see calcserver.hpp, LinkList
        if (e->has("depends"))
            toolAttr(tool, e, "depends");
    } else if (type == "radix-sort" || type == "sort") {
        for (auto a : { "in", "perm", "inv_perm" })
            toolAttr(tool, e, a);
    } else if (type == "unsort") {
        for (auto a : { "in", "out" })
            toolAttr(tool, e, a);
        toolAttr(tool, e, "perm", "id");
    } else if (type == "assert" || type == "if" || type == "while") {
        toolAttr(tool, e, "condition");
    } else if (type == "endif" || type == "end" || type == "dummy") {
    } else if (type == "mpi-sync") {
        toolAttr(tool, e, "mask");
        toolAttr(tool, e, "fields");
        toolAttr(tool, e, "processes", "");
        // (ours, optional) the arrays the mask is a function of: see calcserver.hpp, MPISync
        if (e->has("depends"))
            toolAttr(tool, e, "depends");
    } else if (type == "mpi-allreduce") { // not a reference tool, see calcserver.hpp
        toolAttr(tool, e, "in");
        toolAttr(tool, e, "operation", "min");
    } else if (type == "report_screen") {
        toolAttr(tool, e, "fields");
        toolAttr(tool, e, "bold", "false");
        toolAttr(tool, e, "color", "white");
    } else if (type == "report_file") {
        toolAttr(tool, e, "fields");
        toolAttr(tool, e, "path");
    } else if (type == "report_particles") {
        toolAttr(tool, e, "fields");
        toolAttr(tool, e, "path");
        toolAttr(tool, e, "set");
        toolAttr(tool, e, "ipf", "1");
        toolAttr(tool, e, "fps", "0.0");
    } else if (type == "report_dump") {
        toolAttr(tool, e, "fields");
        toolAttr(tool, e, "path");
        toolAttr(tool, e, "binary", "false");
    } else if (type == "report_performance") {
        toolAttr(tool, e, "bold", "false");
        toolAttr(tool, e, "color", "white");
        toolAttr(tool, e, "path", "");
    } else {
        throw std::runtime_error("Unknown tool type \"" + type + "\" (tool \"" +
                                 tool->get("name") + "\")");
    }
}

void State::parseTools(const Xml::Node* root, ProblemSetup& sd, const std::string& prefix)
{
    for (const Xml::Node* group : root->descendants("Tools"))
        for (const Xml::Node* e : group->descendants("Tool")) {
            if (!e->has("name"))
                throw std::runtime_error("Name shall be defined for tools");
            if (!e->has("type"))
                throw std::runtime_error("Type shall be defined for tools");
            auto tool = std::make_shared<ProblemSetup::Tool>();
            tool->set("name", prefix + e->attr("name"));
            tool->set("type", e->attr("type"));
            tool->set("once", e->has("once") ? toLowerCopy(e->attr("once")) : "false");

            if (e->has("ifdef")) {
                if (!sd.isDefined(e->attr("ifdef")))
                    continue;
            } else if (e->has("ifndef")) {
                if (sd.isDefined(e->attr("ifndef")))
                    continue;
            }

            const std::string action = e->has("action") ? e->attr("action") : "add";
            if (action == "add") {
                sd.tools.push_back(tool);
            } else if (action == "insert" || action == "try_insert") {
                const bool try_insert = action == "try_insert";
                std::vector<unsigned> places;
                bool missing = false;
                if (e->has("at")) {
                    places.push_back(std::stoi(e->attr("at")));
                } else if (e->has("before") || e->has("before_prefix") || e->has("after") ||
                           e->has("after_prefix")) {
                    const bool before = e->has("before") || e->has("before_prefix");
                    std::string att, att_prefix;
                    if (before) {
                        att = e->has("before") ? e->attr("before") : e->attr("before_prefix");
                        att_prefix = e->has("before") ? "" : prefix;
                    } else {
                        att = e->has("after") ? e->attr("after") : e->attr("after_prefix");
                        att_prefix = e->has("after") ? "" : prefix;
                    }
                    const unsigned off = before ? 0 : 1;
                    if (att.find(',') != std::string::npos) {
                        // a list of names: first match (before) / last match (after)
                        auto all = toolsList(att, sd, att_prefix);
                        if (all.empty())
                            missing = true;
                        else
                            places.push_back((before ? *std::min_element(all.begin(), all.end())
                                                     : *std::max_element(all.begin(), all.end())) +
                                             off);
                    } else {
                        // a wildcard: every match
                        auto all = toolsName(att, sd, att_prefix);
                        if (all.empty())
                            missing = true;
                        for (auto p : all)
                            places.push_back(p + off);
                    }
                    if (missing) {
                        if (try_insert)
                            continue;
                        throw std::runtime_error("The tool \"" + tool->get("name") +
                                                 "\" must be inserted relative to \"" + att +
                                                 "\", but such tool cannot be found");
                    }
                } else {
                    throw std::runtime_error("Missing the place where the tool \"" +
                                             tool->get("name") + "\" should be inserted");
                }
                // insert backwards so the earlier places stay valid (State.cpp:917-922)
                for (size_t k = places.size(); k > 0; k--) {
                    const unsigned at = std::min<unsigned>(places[k - 1], sd.tools.size());
                    sd.tools.insert(sd.tools.begin() + at, tool);
                }
            } else if (action == "remove" || action == "try_remove") {
                auto places = toolsName(tool->get("name"), sd, prefix);
                if (places.empty()) {
                    if (action == "try_remove")
                        continue;
                    throw std::runtime_error("Failure removing the tool \"" + tool->get("name") +
                                             "\": no such tool");
                }
                for (size_t k = places.size(); k > 0; k--)
                    sd.tools.erase(sd.tools.begin() + places[k - 1]);
                continue;
            } else if (action == "replace" || action == "try_replace") {
                auto places = toolsName(tool->get("name"), sd, prefix);
                if (places.empty()) {
                    if (action == "try_replace")
                        continue;
                    throw std::runtime_error("Failure replacing the tool \"" + tool->get("name") +
                                             "\": no such tool");
                }
                for (auto p : places)
                    sd.tools[p] = tool;
            } else {
                throw std::runtime_error("Unknown action \"" + action + "\" for the tool \"" +
                                         tool->get("name") + "\"");
            }
            configureTool(tool.get(), e, sd);
        }
}

void State::parseReports(const Xml::Node* root, ProblemSetup& sd, const std::string& prefix)
{
    for (const Xml::Node* group : root->descendants("Reports"))
        for (const Xml::Node* e : group->descendants("Report")) {
            if (!e->has("name") || !e->has("type"))
                throw std::runtime_error("Found a report without name or type");
            auto rep = std::make_shared<ProblemSetup::Tool>();
            rep->set("name", prefix + e->attr("name"));
            const std::string type = e->attr("type");
            rep->set("type", type);
            sd.reports.push_back(rep);
            if (type == "screen") {
                toolAttr(rep.get(), e, "fields");
                toolAttr(rep.get(), e, "bold", "false");
                toolAttr(rep.get(), e, "color", "white");
            } else if (type == "file") {
                toolAttr(rep.get(), e, "fields");
                toolAttr(rep.get(), e, "path");
            } else if (type == "particles") {
                toolAttr(rep.get(), e, "fields");
                toolAttr(rep.get(), e, "path");
                toolAttr(rep.get(), e, "set");
                toolAttr(rep.get(), e, "ipf", "1");
                toolAttr(rep.get(), e, "fps", "0.0");
            } else if (type == "performance") {
                toolAttr(rep.get(), e, "bold", "false");
                toolAttr(rep.get(), e, "color", "white");
                toolAttr(rep.get(), e, "path", "");
            } else {
                throw std::runtime_error("Unknown report type \"" + type + "\"");
            }
        }
}

void State::parseTiming(const Xml::Node* root, ProblemSetup& sd)
{
    typedef ProblemSetup::TimeOpts T;
    for (const Xml::Node* group : root->descendants("Timing"))
        for (const Xml::Node* e : group->descendants("Option")) {
            const std::string name = e->attr("name"), type = e->attr("type");
            if (name == "End" || name == "SimulationStop") {
                if (type == "Time" || type == "T") {
                    sd.time_opts.sim_end_mode |= T::TIME_MODE;
                    sd.time_opts.sim_end_time = std::stof(e->attr("value"));
                } else if (type == "Steps" || type == "S") {
                    sd.time_opts.sim_end_mode |= T::ITER_MODE;
                    sd.time_opts.sim_end_step = std::stoi(e->attr("value"));
                } else if (type == "Frames" || type == "F") {
                    sd.time_opts.sim_end_mode |= T::FRAME_MODE;
                    sd.time_opts.sim_end_frame = std::stoi(e->attr("value"));
                } else {
                    throw std::runtime_error("Unknown simulation stop criteria \"" + type + "\"");
                }
            } else if (name == "Output") {
                if (type == "No") {
                    sd.time_opts.output_mode = T::NO_OUTPUT;
                } else if (type == "FPS") {
                    sd.time_opts.output_mode |= T::FPS_MODE;
                    sd.time_opts.output_fps = std::stof(e->attr("value"));
                } else if (type == "IPF") {
                    sd.time_opts.output_mode |= T::IPF_MODE;
                    sd.time_opts.output_ipf = std::stoi(e->attr("value"));
                } else {
                    throw std::runtime_error("Unknown output criteria \"" + type + "\"");
                }
            } else {
                throw std::runtime_error("Unknown timing option \"" + name + "\"");
            }
        }
}

void State::parseSets(const Xml::Node* root, ProblemSetup& sd)
{
    for (const Xml::Node* e : root->descendants("ParticlesSet")) {
        auto set = std::make_shared<ProblemSetup::ParticlesSet>();
        if (e->has("n")) {
            set->n = (size_t)std::stoll(e->attr("n"));
            set->n_known = true;
        }
        for (const Xml::Node* s : e->descendants("Scalar"))
            set->scalars.emplace_back(s->attr("name"), s->attr("value"));
        for (const Xml::Node* s : e->descendants("Load")) {
            set->in_path = s->attr("file");
            set->in_format = s->attr("format");
            set->in_fields = s->attr("fields");
            if (!_xml_paths.empty() && fs::path(set->in_path).is_relative()) {
                std::error_code ec;
                const fs::path f = fs::path(_xml_paths.back()) / set->in_path;
                if (!fs::exists(set->in_path, ec) && fs::exists(f, ec))
                    set->in_path = f.string();
            }
        }
        for (const Xml::Node* s : e->descendants("Save"))
            set->outputs.push_back({ s->attr("file"), s->attr("format"), s->attr("fields") });
        sd.sets.push_back(set);
    }
}

static std::string scriptRelPath(const std::string& p)
{
    size_t best = std::string::npos;
    for (size_t s = p.find("Scripts/"); s != std::string::npos; s = p.find("Scripts/", s + 8))
        best = s;
    if (best != std::string::npos)
        return p.substr(best);
    return fs::path(p).filename().string();
}

void State::write(const std::string& output_file, const ProblemSetup& sd,
                  bool relative_script_paths, bool full_load_paths) const
{
    using Xml::encodeEntities;
    std::ofstream f(output_file);
    if (!f)
        throw std::runtime_error("Cannot write " + output_file);
    f << "<?xml version=\"1.0\" ?>\n<sphInput>\n";
    f << "    <Settings>\n";
    f << "        <SaveOnFail value=\"" << (sd.settings.save_on_fail ? "true" : "false") << "\" />\n";
    for (auto& d : sd.settings.devices)
        f << "        <Device platform=\"" << d.platform << "\" device=\"" << d.device
          << "\" type=\"" << d.type << "\" addr_bits=\"" << d.addr_bits << "\" />\n";
    f << "    </Settings>\n    <Variables>\n";
    for (auto& v : sd.variables) {
        f << "        <Variable name=\"" << encodeEntities(v.name) << "\" type=\""
          << encodeEntities(v.type) << "\"";
        if (v.type.find('*') != std::string::npos)
            f << " length=\"" << encodeEntities(v.length) << "\"";
        else
            f << " value=\"" << encodeEntities(v.value) << "\"";
        f << " />\n";
    }
    f << "    </Variables>\n    <Definitions>\n";
    for (auto& d : sd.definitions) {
        f << "        <Define name=\"" << encodeEntities(d.name) << "\"";
        if (!d.value.empty() || d.evaluate)
            f << " value=\"" << encodeEntities(d.value) << "\" evaluate=\""
              << (d.evaluate ? "true" : "false") << "\"";
        f << " />\n";
    }
    f << "    </Definitions>\n    <Tools>\n";
    for (auto& t : sd.tools) {
        f << "        <Tool action=\"add\"";
        std::string op;
        for (auto& kv : t->data) {
            if (kv.first == "operation") {
                op = kv.second;
                continue;
            }
            std::string v = kv.second;
            if (kv.first == "path" && relative_script_paths &&
                (t->get("type") == "kernel" || t->get("type") == "python"))
                v = scriptRelPath(v);
            f << " " << kv.first << "=\"" << encodeEntities(v) << "\"";
        }
        if (op.empty())
            f << " />\n";
        else
            f << ">" << encodeEntities(op) << "</Tool>\n";
    }
    f << "    </Tools>\n    <Reports>\n";
    for (auto& t : sd.reports) {
        f << "        <Report";
        for (auto& kv : t->data)
            f << " " << kv.first << "=\"" << encodeEntities(kv.second) << "\"";
        f << " />\n";
    }
    f << "    </Reports>\n    <Timing>\n";
    typedef ProblemSetup::TimeOpts T;
    if (sd.time_opts.sim_end_mode & T::TIME_MODE)
        f << "        <Option name=\"End\" type=\"Time\" value=\"" << sd.time_opts.sim_end_time << "\" />\n";
    if (sd.time_opts.sim_end_mode & T::ITER_MODE)
        f << "        <Option name=\"End\" type=\"Steps\" value=\"" << sd.time_opts.sim_end_step << "\" />\n";
    if (sd.time_opts.sim_end_mode & T::FRAME_MODE)
        f << "        <Option name=\"End\" type=\"Frames\" value=\"" << sd.time_opts.sim_end_frame << "\" />\n";
    if (sd.time_opts.output_mode & T::FPS_MODE)
        f << "        <Option name=\"Output\" type=\"FPS\" value=\"" << sd.time_opts.output_fps << "\" />\n";
    if (sd.time_opts.output_mode & T::IPF_MODE)
        f << "        <Option name=\"Output\" type=\"IPF\" value=\"" << sd.time_opts.output_ipf << "\" />\n";
    f << "    </Timing>\n";
    for (auto& s : sd.sets) {
        f << "    <ParticlesSet";
        if (s->n_known)
            f << " n=\"" << s->n << "\"";
        f << ">\n";
        for (auto& kv : s->scalars)
            f << "        <Scalar name=\"" << kv.first << "\" value=\"" << encodeEntities(kv.second) << "\" />\n";
        if (!s->in_path.empty())
            f << "        <Load format=\"" << s->in_format << "\" file=\""
              << encodeEntities(full_load_paths ? s->in_path : fs::path(s->in_path).filename().string())
              << "\" fields=\""
              << s->in_fields << "\" />\n";
        for (auto& o : s->outputs)
            f << "        <Save format=\"" << o[1] << "\" file=\"" << encodeEntities(o[0])
              << "\" fields=\"" << o[2] << "\" />\n";
        f << "    </ParticlesSet>\n";
    }
    f << "</sphInput>\n";
}

} // namespace InputOutput
} // namespace Aqua
