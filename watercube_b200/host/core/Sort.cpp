// core/Sort.cpp -- see Sort.h.  Reference: src/core/Sort.cpp.
#include "./Sort.h"

#include <cstdio>

namespace core {

Sort::Sort()
    : num_items_(0), num_bins_(0), grid_res_(1), device_(0), bin_size_(1.0f), handle_(nullptr),
      owns_handle_(false) {}

Sort::~Sort() {
    if (owns_handle_ && handle_) wc_destroy(handle_);  // Sort.cpp:7-11 frees its buffers
}

SortRef Sort::numItems(int n) {
    num_items_ = n;
    return shared_from_this();
}

SortRef Sort::gridRes(int r) {
    grid_res_ = r;
    num_bins_ = r * r * r;  // Sort.cpp:21
    return shared_from_this();
}

SortRef Sort::binSize(float s) {
    bin_size_ = s;
    return shared_from_this();
}

SortRef Sort::device(int ordinal) {
    device_ = ordinal;
    return shared_from_this();
}

SortRef Sort::attach(wc_handle* handle) {
    if (owns_handle_ && handle_) wc_destroy(handle_);
    handle_ = handle;
    owns_handle_ = false;
    return shared_from_this();
}

void Sort::prepareBuffers() {
    if (handle_) return;  // attached to a Fluid's handle, or already prepared
    wc_params p;
    util::check(wc_default_params(&p));
    p.num_particles = num_items_;
    p.grid_res = grid_res_;
    // The native layer derives binSize = size / gridRes (Fluid.cpp:208); pick the size that
    // gives back the requested bin size, and a radius that keeps kernelRadius <= binSize.
    p.size = bin_size_ * (float)grid_res_;
    p.particle_radius = bin_size_ * 0.2f;
    p.device = device_;
    util::check(wc_create(&p, &handle_));
    owns_handle_ = true;
    wc_derived d;
    util::check(wc_get_derived(handle_, &d));
    if (d.bin_size != bin_size_)
        util::log("Sort: binSize %.9g is realised as %.9g (size / gridRes)\n", bin_size_,
                  d.bin_size);
}

void Sort::run(Buffer in_particles, Buffer out_particles) {
    if (!handle_) throw Error(WC_ERR_INVALID, "Sort::run before prepareBuffers()");
    // The reference only ever sorts buffer 1 into buffer 2 (Fluid.cpp:347).
    if (in_particles.handle != handle_ || out_particles.handle != handle_ ||
        in_particles.kind != BufferKind::Particles1 || out_particles.kind != BufferKind::Particles2)
        throw Error(WC_ERR_INVALID, "Sort::run: expected (particle buffer 1, particle buffer 2) "
                                    "of the solver this sorter belongs to");
    util::check(wc_sort_only(handle_));
}

void Sort::printGrids() {
    const std::vector<uint32_t> counts = util::getUints(getCountBuffer(), num_bins_);
    const std::vector<uint32_t> offsets = util::getUints(getOffsetBuffer(), num_bins_);
    std::string line = "bins(count, offset):";
    char item[64];
    for (int i = 0; i < num_bins_; i++) {
        std::snprintf(item, sizeof(item), " (%u, %u)", counts[i], offsets[i]);
        line += item;
    }
    util::log("%s\n", line.c_str());
}

}  // namespace core
