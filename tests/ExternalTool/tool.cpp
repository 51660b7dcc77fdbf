// See tool.hpp.  Built by tests/test_installable.py with
//   g++ -std=c++17 -shared -fPIC -Iinclude -Iaquagpusph_b200/host tool.cpp -Laquagpusph_b200 -laquahost -laquacuda
#include "tool.hpp"

#include <vector>

extern "C" Aqua::CalcServer::Tool* create_object(const std::string name, bool once)
{
    return new Aqua::CalcServer::InstallableDemo(name, once);
}

namespace Aqua {
namespace CalcServer {

void InstallableDemo::setup()
{
    _data = variable("plugin_data", true);
    _calls = variable("plugin_calls", false, true);
}

void InstallableDemo::_execute()
{
    // device work through the same C-ABI the built-in tools use
    const size_t n = _data->length();
    std::vector<float> host(n);
    check(aqc_memcpy_d2h(_C->ctx(), host.data(), _data->dptr(), n * sizeof(float), 1));
    for (auto& v : host)
        v *= 2.f;
    check(aqc_memcpy_h2d(_C->ctx(), _data->dptr(), host.data(), n * sizeof(float), 1));
    unsigned calls = 0;
    memcpy(&calls, _calls->get(), sizeof(calls));
    calls++;
    _calls->set(&calls);
}

} // namespace CalcServer
} // namespace Aqua
