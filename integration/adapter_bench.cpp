// Host-side cost of the resampler adapter per call (TEST/TUNING tool, built by `make -C oracle adapter_bench`): where the time of
// newresampler_gpu::metric_resample goes for the sizes newmsm uses (ico6 data grid -> ico4 control grid, D = 1).
#include <chrono>
#include <cstdio>

#include "newmsm_b200/resampler_adapter.hpp"

using namespace newresampler;
using clk = std::chrono::steady_clock;
static double since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

int main(int argc, char** argv) {
    const int hi = argc > 1 ? std::atoi(argv[1]) : 6, lo = argc > 2 ? std::atoi(argv[2]) : 4, D = argc > 3 ? std::atoi(argv[3]) : 1;
    Mesh in = newresampler_gpu::make_mesh_from_icosa(hi), low = newresampler_gpu::make_mesh_from_icosa(lo);
    true_rescale(in, RAD);
    true_rescale(low, RAD);
    in.initialize_pvalues(D);
    for (int d = 0; d < D; ++d)
        for (int v = 0; v < in.nvertices(); ++v) in.set_pvalue(v, 0.01 * in.get_coord(v).X * (d + 1), d);
    for (int rep = 0; rep < 6; ++rep) {
        auto t0 = clk::now();
        Mesh out = newresampler_gpu::metric_resample(in, low, 1);
        const double total = since(t0);
        t0 = clk::now();
        { newresampler_gpu::detail::DeviceMesh a(in), b(low); }
        const double dm = since(t0);
        t0 = clk::now();
        const std::vector<double> fin = newresampler_gpu::detail::pvalues_of(in);
        const double pv = since(t0);
        t0 = clk::now();
        Mesh cp = newresampler_gpu::detail::with_pvalues(low, D, std::vector<double>((size_t)D * low.nvertices(), 0.0));
        const double wp = since(t0);
        newresampler_gpu::detail::DeviceMesh a(in), b(low);
        std::vector<double> fout((size_t)D * low.nvertices());
        t0 = clk::now();
        newresampler_gpu::detail::check(msmgpu_metric_resample(a.h, b.h, D, fin.data(), fout.data()));
        const double core = since(t0);
        t0 = clk::now();
        Mesh sp = low;
        newresampler_gpu::sphere_project_warp(sp, in, in, 1);
        const double warp = since(t0);
        std::printf("rep %d: metric_resample %d->%d D=%d total %.2f ms | 2x DeviceMesh %.2f | pvalues_of %.2f | with_pvalues %.2f | msmgpu_metric_resample %.2f | sphere_project_warp %.2f\n",
                    rep, in.nvertices(), low.nvertices(), D, total, dm, pv, wp, core, warp);
    }
    return 0;
}
