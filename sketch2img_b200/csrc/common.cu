#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace s2i {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------------------------ profiler
bool g_prof_on = false;
long g_alloc_gen = 0;
bool g_prev_kernel = false;
bool g_pdl = [] {
    const char* e = getenv("S2I_NO_PDL");
    return !(e && e[0] == '1');
}();
namespace {
struct ProfRec {
    const char* tag;
    double flops, bytes;
};
cudaStream_t g_prof_stream = nullptr;
std::vector<cudaEvent_t> g_prof_pool;
std::vector<ProfRec> g_prof_recs;
size_t g_prof_used = 0;

cudaEvent_t prof_event() {
    if (g_prof_used == g_prof_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_prof_pool.push_back(e);
    }
    return g_prof_pool[g_prof_used++];
}
}  // namespace

int prof_begin(cudaStream_t st) {
    g_prof_stream = st;
    g_prof_used = 0;
    g_prof_recs.clear();
    cudaEventRecord(prof_event(), st);
    g_prof_on = true;
    return 0;
}

// Roofline denominators for the per-launch floor  max(FLOPs / peak_flops, bytes / peak_bw)  reported by prof_end (0: not set)
static double g_peak_flops = 0.0, g_peak_bw = 0.0;
void prof_set_peaks(double tflops, double gbs) {
    g_peak_flops = tflops * 1e12;
    g_peak_bw = gbs * 1e9;
}

void prof_mark(const char* tag, double flops, double bytes) {
    cudaEventRecord(prof_event(), g_prof_stream);
    g_prof_recs.push_back(ProfRec{tag, flops, bytes});
}

int prof_end(char* buf, int cap) {
    g_prof_on = false;
    if (cudaStreamSynchronize(g_prof_stream) != cudaSuccess)
        return set_error(S2I_ERR_CUDA, "prof_end: %s", cudaGetErrorString(cudaGetLastError()));
    struct Agg {
        long n = 0;
        double ms = 0, flops = 0, bytes = 0, roof_ms = 0;
    };
    std::map<std::string, Agg> agg;
    for (size_t i = 0; i < g_prof_recs.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_prof_pool[i], g_prof_pool[i + 1]);
        Agg& a = agg[g_prof_recs[i].tag];
        a.n += 1;
        a.ms += ms;
        a.flops += g_prof_recs[i].flops;
        a.bytes += g_prof_recs[i].bytes;
        // this launch's roofline floor: whichever of the tensor pipe and HBM bounds it
        double t_f = g_peak_flops > 0 ? g_prof_recs[i].flops / g_peak_flops : 0.0;
        double t_b = g_peak_bw > 0 ? g_prof_recs[i].bytes / g_peak_bw : 0.0;
        a.roof_ms += 1e3 * (t_f > t_b ? t_f : t_b);
    }
    int off = 0;
    for (auto& kv : agg) {
        int w = snprintf(buf + off, cap - off > 0 ? cap - off : 0, "%s %ld %.6f %.6e %.6e %.6f\n", kv.first.c_str(), kv.second.n,
                         kv.second.ms, kv.second.flops, kv.second.bytes, kv.second.roof_ms);
        if (w < 0 || off + w >= cap) return set_error(S2I_ERR_ARG, "prof_end: report buffer too small");
        off += w;
    }
    return off;
}

}  // namespace s2i
