// eb_fft_layout.hpp -- host-only: which tasks of the polyphase FFT field update own a slot of kernel spectra, and how the task array is
// laid out batch by batch (eb_fft.cu::eb_fft_create).  Kept apart from the CUDA translation unit so that tests/tools/eb_layout_check.cpp
// can run it on the CPU.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "eb_fft_core.cuh"

namespace ion {

struct TaskBatch {
    uint32_t c0, nc;  // canonical tasks of the batch: [c0, c0 + nc) of the laid-out array; position inside the batch = slot
    uint32_t m0, nm;  // the mirrored tasks that read those slots
};

// Tasks with 2 ox > dsx read the spectra of the task with x offset dsx - ox (eb_fft_core.cuh, main_phase_product_mirror).
// `tasks` comes from eb_fft_geometry: ox fastest, whole runs ox = 0 .. dsx - 1, so the partner of task i sits at i - ox + (dsx - ox).
// On return mirr[i].kslot is an index into `canon`.
inline void split_mirror_tasks(const std::vector<ebfft::Task>& tasks, uint32_t dsx, std::vector<ebfft::Task>& canon, std::vector<ebfft::Task>& mirr) {
    std::vector<uint32_t> canon_of(tasks.size(), 0xFFFFFFFFu);
    for (size_t i = 0; i < tasks.size(); i++)
        if (2u * tasks[i].ox <= dsx) { canon_of[i] = (uint32_t)canon.size(); canon.push_back(tasks[i]); }
    for (size_t i = 0; i < tasks.size(); i++)
        if (2u * tasks[i].ox > dsx) {
            ebfft::Task t = tasks[i];
            t.mirror = 1u;
            t.kslot = canon_of[i - t.ox + (dsx - t.ox)];
            mirr.push_back(t);
        }
}

// a batch of the mirrored layout holds whole runs of x offsets, so that every mirrored task finds its partner's slot in the same batch
inline uint32_t mirror_batch_slots(uint32_t slots, uint32_t dsx) {
    const uint32_t cpr = dsx / 2u + 1u;  // canonical tasks per run: ox = 0 .. dsx / 2
    slots = slots / cpr * cpr;
    return slots < cpr ? cpr : slots;
}

inline void layout_tasks(const std::vector<ebfft::Task>& tasks, const std::vector<ebfft::Task>& canon, const std::vector<ebfft::Task>& mirr, bool use_mirror,
                         uint32_t batch, std::vector<ebfft::Task>& laid, std::vector<TaskBatch>& batches) {
    laid.reserve(tasks.size());
    if (!use_mirror) {
        for (size_t c0 = 0; c0 < tasks.size(); c0 += batch) {
            const uint32_t nc = (uint32_t)(tasks.size() - c0 < batch ? tasks.size() - c0 : batch);
            for (uint32_t i = 0; i < nc; i++) { ebfft::Task t = tasks[c0 + i]; t.kslot = i; t.mirror = 0u; laid.push_back(t); }
            batches.push_back(TaskBatch{(uint32_t)c0, nc, (uint32_t)c0 + nc, 0u});
        }
        return;
    }
    size_t mi = 0;  // mirr is ordered like canon run by run (both follow the task order), so a batch's mirrored tasks are a contiguous range
    for (size_t c0 = 0; c0 < canon.size(); c0 += batch) {
        const uint32_t nc = (uint32_t)(canon.size() - c0 < batch ? canon.size() - c0 : batch);
        TaskBatch b;
        b.c0 = (uint32_t)laid.size(); b.nc = nc;
        for (uint32_t i = 0; i < nc; i++) { ebfft::Task t = canon[c0 + i]; t.kslot = i; t.mirror = 0u; laid.push_back(t); }
        b.m0 = (uint32_t)laid.size(); b.nm = 0u;
        while (mi < mirr.size() && mirr[mi].kslot < c0 + nc) {
            ebfft::Task t = mirr[mi++];
            t.kslot -= (uint32_t)c0;
            laid.push_back(t);
            b.nm++;
        }
        batches.push_back(b);
    }
}

}  // namespace ion
