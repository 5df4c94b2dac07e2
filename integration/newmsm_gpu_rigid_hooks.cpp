// newmsm_gpu_rigid_hooks.cpp — link-time binding of the AFFINE / RIGID level into the UNMODIFIED reference `newmsm` program.
//
// Rigid_cost_function::run (rigid_costfunction.cpp:167-236) calls rigid_cost_mesh from inside its own translation unit, so — like the
// groupwise members (newmsm_gpu_group_hooks.cpp) — the two members below are made weak in the compiled reference OBJECT with objcopy
// (`make -C oracle newmsm_gpu`) and get a `__real_` alias; the strong definitions in this file win at link time:
//
//   Rigid_cost_function::initialise        -> msmgpu_rigid_create   (TARGET octree, similarity means, neighbourhood emptiness on the device)
//   Rigid_cost_function::rigid_cost_mesh   -> msmgpu_rigid_cost     (csrc/rigid.cu)
//
// `run`, `rotate_in_mesh` and `set_parameters` stay the reference's code. Compiled with -fno-access-control (the binding reads the
// object's private meshes and feature space; a maintainer would put the two calls into the members themselves).
// MSMGPU_DISABLE=rigid keeps the reference's code; MSMGPU_VERIFY=1 evaluates every cost on both paths and compares the bits.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include "NewMeshReg/rigid_costfunction.h"
#include "newmsm_b200/resampler_adapter.hpp"

using newmeshreg::Rigid_cost_function;

#define SYM_RINIT "_ZN10newmeshreg19Rigid_cost_function10initialiseEv"
#define SYM_RCOST "_ZN10newmeshreg19Rigid_cost_function15rigid_cost_meshEddd"

void real_rigid_initialise(Rigid_cost_function* self) asm("__real_" SYM_RINIT);
void hook_rigid_initialise(Rigid_cost_function* self) asm(SYM_RINIT);
double real_rigid_cost_mesh(Rigid_cost_function* self, double dw1, double dw2, double dw3) asm("__real_" SYM_RCOST);
double hook_rigid_cost_mesh(Rigid_cost_function* self, double dw1, double dw2, double dw3) asm(SYM_RCOST);

namespace {

bool disabled() {
    static const bool d = [] { const char* e = std::getenv("MSMGPU_DISABLE"); return e && std::strstr(e, "rigid"); }();
    return d;
}
bool verify() { static const bool v = std::getenv("MSMGPU_VERIFY") != nullptr; return v; }

std::mutex g_mutex;
std::map<Rigid_cost_function*, msmgpu_rigid*> g_state;    // device state of the live cost functions (one per AFFINE level)

struct Report {
    double seconds = 0;
    long evaluations = 0, checked = 0, bad = 0;
    ~Report() {
        for (auto& kv : g_state) msmgpu_rigid_destroy(kv.second);
        if (verify()) std::fprintf(stderr, "[msmgpu verify] rigid_cost_mesh: %ld of %ld costs differ from the reference's\n", bad, checked);
        if (std::getenv("MSMGPU_TIMING") && evaluations)
            std::fprintf(stderr, "[msmgpu rigid] %ld cost evaluations in %.3f s\n", evaluations, seconds);
    }
} report;

std::vector<int32_t> triangles_of(const newresampler::Mesh& m) {
    std::vector<int32_t> tri(3 * (size_t)m.ntriangles());
    for (int t = 0; t < m.ntriangles(); ++t)
        for (int k = 0; k < 3; ++k) tri[3 * (size_t)t + k] = m.get_triangle_vertexID(t, k);
    return tri;
}

std::vector<double> rows_of(const std::shared_ptr<MISCMATHS::BFMatrix>& M) {   // channel-major [D][V] from the 1-based feature matrix
    const int D = (int)M->Nrows(), V = (int)M->Ncols();
    std::vector<double> f((size_t)D * V);
    for (int d = 0; d < D; ++d)
        for (int v = 0; v < V; ++v) f[(size_t)d * V + v] = M->Peek(d + 1, v + 1);
    return f;
}

}  // namespace

void hook_rigid_initialise(Rigid_cost_function* self) {
    if (verify() || disabled()) real_rigid_initialise(self);     // verify: the reference's own state too, for side-by-side costs
    if (disabled()) return;
    using newresampler_gpu::detail::check;
    self->min_sigma = self->MVD = self->SOURCE.calculate_MeanVD();                  // cpp:35
    const std::vector<double> tx = newresampler_gpu::detail::coords_of(self->TARGET), sx = newresampler_gpu::detail::coords_of(self->SOURCE);
    const std::vector<int32_t> tt = triangles_of(self->TARGET), st = triangles_of(self->SOURCE);
    const std::vector<double> A = rows_of(self->FEAT->get_input_data()), B = rows_of(self->FEAT->get_reference_data());
    const int D = (int)self->FEAT->get_input_data()->Nrows();
    msmgpu_rigid* r = nullptr;
    check(msmgpu_rigid_create(newresampler_gpu::detail::context(), self->TARGET.nvertices(), tx.data(), self->TARGET.ntriangles(), tt.data(),
                              self->SOURCE.nvertices(), sx.data(), self->SOURCE.ntriangles(), st.data(), D, A.data(), B.data(), self->simmeasure,
                              self->MVD, &r));
    std::lock_guard<std::mutex> g(g_mutex);
    auto it = g_state.find(self);
    if (it != g_state.end()) msmgpu_rigid_destroy(it->second);
    g_state[self] = r;
}

double hook_rigid_cost_mesh(Rigid_cost_function* self, double dw1, double dw2, double dw3) {
    if (disabled()) return real_rigid_cost_mesh(self, dw1, dw2, dw3);
    msmgpu_rigid* r = nullptr;
    {
        std::lock_guard<std::mutex> g(g_mutex);
        auto it = g_state.find(self);
        if (it != g_state.end()) r = it->second;
    }
    if (!r) return real_rigid_cost_mesh(self, dw1, dw2, dw3);     // initialise() was not called through the hook
    const double t0 = omp_get_wtime();
    const std::vector<double> sx = newresampler_gpu::detail::coords_of(self->SOURCE);
    double cost = 0.0;
    newresampler_gpu::detail::check(msmgpu_rigid_cost(r, sx.data(), dw1, dw2, dw3, &cost));
    report.seconds += omp_get_wtime() - t0;
    ++report.evaluations;
    if (verify()) {
        const double ref = real_rigid_cost_mesh(self, dw1, dw2, dw3);
        ++report.checked;
        if (std::memcmp(&ref, &cost, sizeof(double)) != 0) {
            ++report.bad;
            std::fprintf(stderr, "[msmgpu verify] rigid_cost_mesh(%g, %g, %g): device %.17g, reference %.17g\n", dw1, dw2, dw3, cost, ref);
        }
    }
    return cost;
}
