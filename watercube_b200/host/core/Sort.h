// core/Sort.h -- host facade of the cell-hash binning + stable counting sort.
//
// Same public surface as the reference's core::Sort (src/core/Sort.h:20-39): fluent
// numItems / gridRes / binSize, prepareBuffers, compileShaders, run(in, out) and the three
// buffer accessors.  The three GLSL dispatches of Sort::run (Sort.cpp:254-267) are one
// wc_sort_only() call on the native handle; the buffers live in that handle.
#pragma once

#include <memory>

#include "./util.h"

namespace core {

typedef std::shared_ptr<class Sort> SortRef;

class Sort : public std::enable_shared_from_this<Sort> {
public:
    Sort();
    ~Sort();
    Sort(const Sort&) = delete;
    Sort& operator=(const Sort&) = delete;

    // The reference's setters return a COPY wrapped in a new shared_ptr (Sort.h:56, quirk
    // Q17); here they return the object itself, which is what every chained call site
    // (Fluid.cpp:229) relies on.
    SortRef numItems(int n);
    SortRef gridRes(int r);
    SortRef binSize(float s);
    SortRef device(int ordinal);

    // Use the buffers of an existing solver handle (how Fluid::setup wires its sorter);
    // prepareBuffers() is then a no-op.
    SortRef attach(wc_handle* handle);

    void prepareBuffers();   // Sort.cpp:67-94: allocates count / offset / sorted (+ particles)
    void compileShaders() {} // Sort.cpp:99-130: nothing to compile, kernels are in the library
    void run(Buffer in_particles, Buffer out_particles);  // Sort.cpp:254-267
    void renderGrid(float /*size*/) {}                    // graphics, out of scope

    Buffer getCountBuffer() const { return Buffer{handle_, BufferKind::Counts}; }
    Buffer getOffsetBuffer() const { return Buffer{handle_, BufferKind::Offsets}; }
    Buffer getSortedBuffer() const { return Buffer{handle_, BufferKind::Sorted}; }
    // Particle buffers of a stand-alone sorter (a Fluid hands out its own).
    Buffer getInBuffer() const { return Buffer{handle_, BufferKind::Particles1}; }
    Buffer getOutBuffer() const { return Buffer{handle_, BufferKind::Particles2}; }

    void printGrids();  // Sort.cpp:237-249

    static SortRef create() { return std::make_shared<Sort>(); }

protected:
    int num_items_, num_bins_, grid_res_, device_;
    float bin_size_;
    wc_handle* handle_;
    bool owns_handle_;
};

}  // namespace core
