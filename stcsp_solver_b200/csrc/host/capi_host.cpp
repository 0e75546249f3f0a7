// C entry points of the host side (include/stcsp_host.h) and the shared error slot.
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "error.h"
#include "model.h"
#include "stcsp_host.h"

namespace stcsp {
namespace {
thread_local std::string g_last_error;
}
void set_error(const std::string &msg) { g_last_error = msg; }
}  // namespace stcsp

struct stcsp_model {
    stcsp::Model model;
    std::unique_ptr<stcsp::FlatProblem> flat;
};

extern "C" {

const char *stcsp_last_error(void) { return stcsp::g_last_error.c_str(); }

int stcsp_model_parse_text(const char *text, int32_t prefix_k, stcsp_model_t **out) {
    if (!text || !out) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    *out = nullptr;
    try {
        auto m = std::make_unique<stcsp_model>();
        m->model = stcsp::parse_model(text, prefix_k);
        m->flat = stcsp::flatten(m->model);
        *out = m.release();
        return STCSP_OK;
    } catch (const stcsp::ParseError &e) {
        stcsp::set_error(e.what());
        return STCSP_ERR_PARSE;
    } catch (const std::exception &e) {
        stcsp::set_error(e.what());
        return STCSP_ERR_INVALID;
    }
}

int stcsp_model_parse_file(const char *path, int32_t prefix_k, stcsp_model_t **out) {
    if (!path || !out) { stcsp::set_error("null argument"); return STCSP_ERR_INVALID; }
    std::ifstream in(path, std::ios::binary);
    if (!in) { stcsp::set_error(std::string("cannot open ") + path); return STCSP_ERR_PARSE; }
    std::stringstream ss;
    ss << in.rdbuf();
    return stcsp_model_parse_text(ss.str().c_str(), prefix_k, out);
}

void stcsp_model_free(stcsp_model_t *m) { delete m; }

const stcsp_problem_t *stcsp_model_problem(const stcsp_model_t *m) { return m ? &m->flat->c : nullptr; }

char *stcsp_model_dump(const stcsp_model_t *m) {
    std::string s = stcsp::dump_model(m->model);
    char *p = (char *)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

}  // extern "C"
