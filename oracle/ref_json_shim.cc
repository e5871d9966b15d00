// ref_json_shim.cc — builds oracle/_ref/libref_json.so around the REFERENCE's own src/json.h
// (header-only, compiled from where it lies under /root/reference; never copied into this repo).
// It assembles the result object exactly as BatchRecognizer::PushLattice does
// [REF src/batch_recognizer.cc:82-105] and returns json::JSON::dump() [REF src/json.h:343-384],
// pinning the result-text layout of the oracle and of the engine.  TEST INFRASTRUCTURE ONLY.
#include "json.h"

#include <cstdlib>
#include <cstring>

extern "C" char *ref_json_result(int n, const char *const *words, const double *start, const double *end,
                                 const double *conf, const char *text) {
    json::JSON obj;
    for (int i = 0; i < n; i++) {
        json::JSON word;
        word["word"] = words[i];
        word["start"] = start[i];
        word["end"] = end[i];
        word["conf"] = conf[i];
        obj["result"].append(word);
    }
    obj["text"] = text;
    std::string s = obj.dump();
    char *r = (char *)malloc(s.size() + 1);
    memcpy(r, s.c_str(), s.size() + 1);
    return r;
}
extern "C" void ref_json_free(char *p) { free(p); }
