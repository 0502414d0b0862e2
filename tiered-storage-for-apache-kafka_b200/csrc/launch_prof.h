// launch_prof.h — launch counting and optional per-kernel CUDA-event timing (bench.py's live roofline numbers).
#pragma once
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>
#include "rt.h"

namespace ts {

struct LaunchProf {
    std::atomic<uint64_t> launches{0};      // calls on one context run concurrently (slot ownership, tsgpu.cu)
    bool on = false;
    struct Rec { const char* name; rt::tevent_t a, b; };
    std::vector<Rec> recs;
    std::mutex m;
    // returns the record's index (SIZE_MAX when timing is off) for end()
    size_t begin(const char* name, rt::stream_t st) {
        launches++;
        if (!on) return (size_t)-1;
        Rec r{name, {}, {}};
        rt::tevent_create(&r.a); rt::tevent_create(&r.b);
        rt::tevent_record(r.a, st);
        std::lock_guard<std::mutex> lock(m);
        recs.push_back(r);
        return recs.size() - 1;
    }
    void end(size_t i, rt::stream_t st) {
        if (i == (size_t)-1) return;
        std::lock_guard<std::mutex> lock(m);
        if (i < recs.size()) rt::tevent_record(recs[i].b, st);
    }
    void reset() {
        std::lock_guard<std::mutex> lock(m);
        for (auto& r : recs) { rt::tevent_destroy(r.a); rt::tevent_destroy(r.b); }
        recs.clear();
    }
    // caller must have synchronised the streams
    std::string report_json() {
        std::map<std::string, std::pair<uint64_t, double>> acc;
        for (auto& r : recs) { auto& e = acc[r.name]; e.first++; e.second += rt::tevent_ms(r.a, r.b); }
        std::string s = "{";
        bool first = true;
        for (auto& kv : acc) {
            char buf[256];
            snprintf(buf, sizeof buf, "%s\"%s\":{\"launches\":%llu,\"ms\":%.6f}", first ? "" : ",", kv.first.c_str(),
                     (unsigned long long)kv.second.first, kv.second.second);
            s += buf; first = false;
        }
        return s + "}";
    }
};

}  // namespace ts

#define TS_LAUNCH_P(prof, name, kern, grid, block, smem, stream, ...) \
    do { const size_t pi_ = (prof).begin(name, stream); TS_LAUNCH(kern, grid, block, smem, stream, __VA_ARGS__); (prof).end(pi_, stream); } while (0)
