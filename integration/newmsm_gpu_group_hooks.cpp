// newmsm_gpu_group_hooks.cpp — link-time binding of the groupwise (gMSM) paths into the UNMODIFIED reference `newmsm` program.
//
// The groupwise model calls its hot members from inside its own translation unit (setupCostFunction -> estimate_pairs,
// get_patch_data; DiscreteGroupModel.cpp:166-195) and its cost function is reached through the vtable, so the linker's --wrap
// (which only redirects UNDEFINED references) cannot see them. Instead `make -C oracle newmsm_gpu` post-processes the compiled
// reference OBJECTS (never the sources) with objcopy: the four symbols below are made weak and get a `__real_` alias at the same
// address; the strong definitions in this file then win at link time, for same-object calls and vtable slots alike, and can still
// fall back to the original code through the alias.
//
//   DiscreteGroupModel::estimate_pairs / get_patch_data            -> newmeshreg_gpu::GroupBinding (include/newmsm_b200/group_adapter.hpp)
//   DiscreteGroupCostFunction::computePairwiseCost / computeTripletCost
//
// Compiled with -fno-access-control: the binding reads the private state of the two reference classes (a maintainer would put
// the same four calls into the members themselves). MSMGPU_DISABLE=group keeps the reference's code;
// MSMGPU_TIMING=1 prints the split at exit.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "newmsm_b200/group_adapter.hpp"

using newmeshreg::DiscreteGroupCostFunction;
using newmeshreg::DiscreteGroupModel;
using newmeshreg_gpu::GroupBinding;

#define SYM_PAIRS "_ZN10newmeshreg18DiscreteGroupModel14estimate_pairsEv"
#define SYM_PATCH "_ZN10newmeshreg18DiscreteGroupModel14get_patch_dataEv"
#define SYM_PCOST "_ZN10newmeshreg25DiscreteGroupCostFunction19computePairwiseCostEiii"
#define SYM_TCOST "_ZN10newmeshreg25DiscreteGroupCostFunction18computeTripletCostEiiii"

void real_estimate_pairs(DiscreteGroupModel* self) asm("__real_" SYM_PAIRS);
void hook_estimate_pairs(DiscreteGroupModel* self) asm(SYM_PAIRS);
void real_get_patch_data(DiscreteGroupModel* self) asm("__real_" SYM_PATCH);
void hook_get_patch_data(DiscreteGroupModel* self) asm(SYM_PATCH);
double real_pair_cost(DiscreteGroupCostFunction* self, int pair, int la, int lb) asm("__real_" SYM_PCOST);
double hook_pair_cost(DiscreteGroupCostFunction* self, int pair, int la, int lb) asm(SYM_PCOST);
double real_triplet_cost(DiscreteGroupCostFunction* self, int t, int la, int lb, int lc) asm("__real_" SYM_TCOST);
double hook_triplet_cost(DiscreteGroupCostFunction* self, int t, int la, int lb, int lc) asm(SYM_TCOST);

namespace {

bool disabled() {
    static const bool d = [] { const char* e = std::getenv("MSMGPU_DISABLE"); return e && std::strstr(e, "group"); }();
    return d;
}

bool verify() { static const bool v = std::getenv("MSMGPU_VERIFY") != nullptr; return v; }
std::atomic<long> g_pair_checked{0}, g_pair_bad{0}, g_trip_checked{0}, g_trip_bad{0};

struct Report {
    ~Report() {
        if (verify())
            std::fprintf(stderr, "[msmgpu verify] group pair costs: %ld of %ld differ from the reference's | group triplet costs: %ld of %ld differ\n",
                         g_pair_bad.load(), g_pair_checked.load(), g_trip_bad.load(), g_trip_checked.load());
        if (!std::getenv("MSMGPU_TIMING")) return;
        const auto& t = newmeshreg_gpu::group_timers();
        if (t.n_iterations == 0) return;
        std::fprintf(stderr,
                     "[msmgpu group] iterations %ld | estimate_pairs %.3f s | get_patch_data (fields + patch geometry) %.3f s | pair batches %ld in %.3f s "
                     "(%lld costs) | triplet batches %ld in %.3f s (%lld costs)\n",
                     t.n_iterations, t.pairs, t.fields, t.n_pair_batches, t.pair_batches, t.pair_costs, t.n_triplet_batches, t.triplet_batches, t.triplet_costs);
    }
} report;

}  // namespace

// a call that leaves the device path says so once (never silently): MSMGPU_DISABLE=group, or a shape the binding does not take
static void left_to_reference(const char* what, bool by_request) {
    static std::atomic<int> said{0};
    if (said.fetch_add(1) < 4)
        std::fprintf(stderr, "[msmgpu] %s runs on the reference's CPU code (%s)\n", what,
                     by_request ? "MSMGPU_DISABLE=group" : "fewer than 2 subjects, or the subjects' data meshes differ in size");
}

void hook_estimate_pairs(DiscreteGroupModel* self) {
    if (disabled() || !GroupBinding::instance().estimate_pairs(*self)) { left_to_reference("DiscreteGroupModel::estimate_pairs", disabled()); real_estimate_pairs(self); return; }
    if (verify()) {
        const std::vector<int> got(self->pairs, self->pairs + 2 * (size_t)self->m_num_pairs);
        real_estimate_pairs(self);
        long bad = 0;
        for (size_t i = 0; i < got.size(); ++i) bad += got[i] != self->pairs[i];
        std::fprintf(stderr, "[msmgpu verify] estimate_pairs: %d pairs, %ld node ids differ from the reference's\n", self->m_num_pairs, bad);
    }
}

void hook_get_patch_data(DiscreteGroupModel* self) {
    if (disabled() || !GroupBinding::instance().get_patch_data(*self)) { left_to_reference("DiscreteGroupModel::get_patch_data and the group costs", disabled()); real_get_patch_data(self); return; }
    if (verify()) real_get_patch_data(self);   // the reference's patch maps too, for the side-by-side comparison of every cost request
    // Fusion::optimize runs next; the first sphere_project_warp after it closes the "optimiser phases" timer (newmsm_gpu_hooks.cpp)
    newmeshreg_gpu::detail::timers().source_done_at = omp_get_wtime();
}

double hook_pair_cost(DiscreteGroupCostFunction* self, int pair, int la, int lb) {
    GroupBinding& g = GroupBinding::instance();
    if (!g.active_for(self)) return real_pair_cost(self, pair, la, lb);
    const double v = g.pairwise(*self, pair, la, lb);
    if (verify()) {
        const double r = real_pair_cost(self, pair, la, lb);
        g_pair_checked++;
        if (std::memcmp(&r, &v, sizeof(double)) != 0 && !(std::isnan(r) && std::isnan(v)) && g_pair_bad++ < 20)
            std::fprintf(stderr, "[msmgpu verify] group pair %d labels %d %d: reference %a (%.17g)  device %a (%.17g)\n", pair, la, lb, r, r, v, v);
    }
    return v;
}

double hook_triplet_cost(DiscreteGroupCostFunction* self, int t, int la, int lb, int lc) {
    GroupBinding& g = GroupBinding::instance();
    if (!g.active_for(self)) return real_triplet_cost(self, t, la, lb, lc);
    const double v = g.triplet(*self, t, la, lb, lc);
    if (verify()) {
        const double r = real_triplet_cost(self, t, la, lb, lc);
        g_trip_checked++;
        if (std::memcmp(&r, &v, sizeof(double)) != 0 && !(std::isnan(r) && std::isnan(v)) && g_trip_bad++ < 20)
            std::fprintf(stderr, "[msmgpu verify] group triplet %d labels %d %d %d: reference %a (%.17g)  device %a (%.17g)\n", t, la, lb, lc, r, r, v, v);
    }
    return v;
}
