// CPU check of ionsolver_b200/csrc/eb_fft_layout.hpp (which tasks of the FFT field update own kernel spectra, batch layout).
// For many block shapes (dsx, dsy, dsz, extra z window) and slot capacities: every task appears exactly once; a canonical task's slot is
// its position inside its batch; a mirrored task points at the canonical task with x offset dsx - ox and equal (oy, oz, wz) INSIDE ITS
// OWN BATCH; batches respect the capacity; the share of canonical tasks is (dsx / 2 + 1) / dsx.
// usage: eb_layout_check   -> one JSON line, exit code 0 when every invariant holds
#include <cstdio>
#include <map>
#include <tuple>
#include <vector>

#include "../../ionsolver_b200/csrc/eb_fft_layout.hpp"

using ion::ebfft::Task;

int main() {
    long cases = 0, bad = 0;
    for (uint32_t dsx : {1u, 2u, 3u, 4u, 5u, 8u, 16u})
        for (uint32_t dsy : {1u, 3u})
            for (uint32_t dsz : {1u, 2u})
                for (uint32_t nwz : {1u, 2u}) {
                    std::vector<Task> tasks;  // the order of eb_fft_geometry: wz, oz, oy, ox (ox fastest)
                    for (uint32_t wz = 0; wz < nwz; wz++)
                        for (uint32_t oz = 0; oz < dsz; oz++) {
                            if (wz == 1u && oz > 0u) continue;  // a partial extra window, like a slab with two halo layers
                            for (uint32_t oy = 0; oy < dsy; oy++)
                                for (uint32_t ox = 0; ox < dsx; ox++) tasks.push_back(Task{(uint16_t)ox, (uint16_t)oy, (uint16_t)oz, (uint16_t)wz, 0u, 0u});
                        }
                    std::vector<Task> canon, mirr;
                    ion::split_mirror_tasks(tasks, dsx, canon, mirr);
                    if (canon.size() + mirr.size() != tasks.size()) bad++;
                    if (canon.size() * dsx != tasks.size() * (dsx / 2u + 1u)) bad++;
                    for (int use_mirror = 0; use_mirror < 2; use_mirror++) {
                        if (use_mirror && mirr.empty()) continue;
                        for (uint32_t cap : {1u, 2u, 3u, 5u, 7u, 9u, 64u, 100000u}) {
                            const uint32_t all = use_mirror ? (uint32_t)canon.size() : (uint32_t)tasks.size();
                            uint32_t batch = cap < all ? cap : all;
                            if (use_mirror && batch < all) batch = ion::mirror_batch_slots(batch, dsx);
                            std::vector<Task> laid;
                            std::vector<ion::TaskBatch> batches;
                            ion::layout_tasks(tasks, canon, mirr, use_mirror != 0, batch, laid, batches);
                            cases++;
                            if (laid.size() != tasks.size()) { bad++; continue; }
                            std::map<std::tuple<int, int, int, int>, int> seen;
                            size_t covered = 0;
                            for (const ion::TaskBatch& b : batches) {
                                if (b.nc > batch || b.nc == 0u || b.m0 != b.c0 + b.nc || b.c0 != covered) bad++;
                                covered += b.nc + b.nm;
                                for (uint32_t i = 0; i < b.nc; i++) {
                                    const Task& t = laid[b.c0 + i];
                                    if (t.kslot != i || t.mirror != 0u || (use_mirror && 2u * t.ox > dsx)) bad++;
                                    seen[std::make_tuple(t.ox, t.oy, t.oz, t.wz)]++;
                                }
                                for (uint32_t i = 0; i < b.nm; i++) {
                                    const Task& t = laid[b.m0 + i];
                                    if (t.mirror != 1u || t.kslot >= b.nc || 2u * t.ox <= dsx) { bad++; continue; }
                                    const Task& c = laid[b.c0 + t.kslot];
                                    if (c.ox != dsx - t.ox || c.oy != t.oy || c.oz != t.oz || c.wz != t.wz) bad++;
                                    seen[std::make_tuple(t.ox, t.oy, t.oz, t.wz)]++;
                                }
                            }
                            if (covered != tasks.size() || seen.size() != tasks.size()) bad++;
                            for (const auto& kv : seen)
                                if (kv.second != 1) bad++;
                        }
                    }
                }
    printf("{\"layouts_checked\": %ld, \"violations\": %ld}\n", cases, bad);
    return bad == 0 ? 0 : 1;
}
