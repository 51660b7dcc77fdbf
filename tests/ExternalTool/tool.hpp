// A type="installable" tool for THIS host: the counterpart of the reference's demo plugin
// (tests/ExternalTool/tool.hpp:28-40, tool.cpp:27-32 there): a shared library that exports
//     Tool* create_object(const std::string name, bool once)
// and whose class derives from Aqua::CalcServer::Tool.  Unlike the reference's demo, which only
// logs a line, this one does observable work: every execution doubles the device array
// `plugin_data` through the C-ABI and counts itself in the scalar `plugin_calls`.
#pragma once
#include "calcserver.hpp"

extern "C" {
Aqua::CalcServer::Tool* create_object(const std::string name, bool once);
}

namespace Aqua {
namespace CalcServer {

class InstallableDemo : public Tool {
  public:
    InstallableDemo(const std::string name, bool once) : Tool(name, once) {}
    void setup() override;

  protected:
    void _execute() override;

  private:
    InputOutput::Variable *_data = nullptr, *_calls = nullptr;
};

} // namespace CalcServer
} // namespace Aqua
