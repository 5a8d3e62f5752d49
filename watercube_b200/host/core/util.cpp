// core/util.cpp -- see util.h.  Reference: src/core/util.cpp.
#include "./util.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace core {
namespace util {

void check(int rc) {
    if (rc != WC_OK) throw Error(rc, std::string("wc_sph: ") + wc_last_error());
}

void log(const char* format, ...) {
    va_list args;
    va_start(args, format);
    std::vfprintf(stderr, format, args);
    va_end(args);
}

static wc_handle* need(Buffer b, const char* who) {
    if (!b) throw Error(WC_ERR_INVALID, std::string(who) + ": null buffer");
    return b.handle;
}

std::vector<Particle> getParticles(Buffer buffer, int num_items) {
    wc_handle* h = need(buffer, "getParticles");
    if (buffer.kind != BufferKind::Particles1 && buffer.kind != BufferKind::Particles2)
        throw Error(WC_ERR_INVALID, "getParticles: not a particle buffer");
    if (buffer.slabs) {  // decomposed fluid: the slabs in z order
        std::vector<Particle> all;
        for (wc_handle* s : *buffer.slabs) {
            int32_t count = 0;
            check(wc_get_num_particles(s, &count));  // (waits for the slab's queued steps)
            const size_t at = all.size();
            all.resize(at + (size_t)count);
            check(wc_download_particles(s, (int)buffer.kind, reinterpret_cast<wc_particle*>(all.data() + at)));
        }
        if (num_items >= 0 && (size_t)num_items < all.size()) all.resize((size_t)num_items);
        return all;
    }
    int32_t count = 0;
    check(wc_get_num_particles(h, &count));
    std::vector<Particle> all((size_t)count);
    check(wc_download_particles(h, (int)buffer.kind, reinterpret_cast<wc_particle*>(all.data())));
    if (num_items >= 0 && (size_t)num_items < all.size()) all.resize((size_t)num_items);
    return all;
}

void setParticles(Buffer buffer, const std::vector<Particle>& particles) {
    wc_handle* h = need(buffer, "setParticles");
    if (buffer.slabs)
        throw Error(WC_ERR_INVALID, "setParticles: a decomposed fluid takes new particles through "
                                    "Fluid::initialParticles (they must be cut into slabs)");
    const wc_particle* src = reinterpret_cast<const wc_particle*>(particles.data());
    if (buffer.kind == BufferKind::Particles1)
        check(wc_upload_particles(h, src, (int32_t)particles.size()));
    else if (buffer.kind == BufferKind::Particles2)
        check(wc_upload_sorted(h, src, (int32_t)particles.size()));
    else
        throw Error(WC_ERR_INVALID, "setParticles: not a particle buffer");
}

std::vector<uint32_t> getUints(Buffer buffer, int num_items) {
    wc_handle* h = need(buffer, "getUints");
    if (buffer.slabs)
        throw Error(WC_ERR_INVALID, "getUints: the cell tables of a decomposed fluid are per slab "
                                    "(Fluid::slabHandles + wc_download_cells)");
    int32_t count = 0;
    check(wc_get_num_particles(h, &count));
    wc_derived d;
    check(wc_get_derived(h, &d));
    const bool per_bin = buffer.kind == BufferKind::Counts || buffer.kind == BufferKind::Offsets;
    std::vector<uint32_t> out(per_bin ? (size_t)d.num_bins : (size_t)count);
    uint32_t* p = out.data();
    switch (buffer.kind) {
        case BufferKind::CellIds: check(wc_download_cells(h, p, nullptr, nullptr, nullptr, nullptr)); break;
        case BufferKind::Counts: check(wc_download_cells(h, nullptr, p, nullptr, nullptr, nullptr)); break;
        case BufferKind::Offsets: check(wc_download_cells(h, nullptr, nullptr, p, nullptr, nullptr)); break;
        case BufferKind::Sorted: check(wc_download_cells(h, nullptr, nullptr, nullptr, p, nullptr)); break;
        case BufferKind::NeighbourCounts:
            check(wc_download_cells(h, nullptr, nullptr, nullptr, nullptr, p));
            break;
        default: throw Error(WC_ERR_INVALID, "getUints: not a uint buffer");
    }
    if (num_items >= 0 && (size_t)num_items < out.size()) out.resize((size_t)num_items);
    return out;
}

void printParticles(Buffer particle_buffer, int n, float bin_size) {
    const std::vector<Particle> particles = getParticles(particle_buffer, n);
    for (size_t i = 0; i < particles.size(); i++) {
        const Particle& p = particles[i];
        log("p=<%f, %f, %f>, v=<%f, %f, %f>, d=%f, pr=%f, c=<%d, %d, %d>\n", p.position.x,
            p.position.y, p.position.z, p.velocity.x, p.velocity.y, p.velocity.z, p.density,
            p.pressure, (int)(p.position.x / bin_size), (int)(p.position.y / bin_size),
            (int)(p.position.z / bin_size));
    }
}

static const char kMagic[8] = {'W', 'C', 'B', '2', '0', '0', 0, 0};

void saveCheckpoint(const std::string& path, const CheckpointHeader& header,
                    const std::vector<Particle>& particles) {
    CheckpointHeader h = header;
    std::memcpy(h.magic, kMagic, sizeof(kMagic));
    h.version = 2;
    h.num_particles = (int32_t)particles.size();
    FILE* f = std::fopen(path.c_str(), "wb");
    const bool ok = f && std::fwrite(&h, sizeof(h), 1, f) == 1 &&
                    std::fwrite(particles.data(), sizeof(Particle), particles.size(), f) ==
                        particles.size();
    if (f && std::fclose(f) != 0) throw Error(WC_ERR_INVALID, "saveCheckpoint: cannot finish " + path);
    if (!ok) throw Error(WC_ERR_INVALID, "saveCheckpoint: cannot write " + path);
}

std::vector<Particle> loadCheckpoint(const std::string& path, CheckpointHeader* header) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) throw Error(WC_ERR_INVALID, "loadCheckpoint: cannot open " + path);
    CheckpointHeader h = CheckpointHeader();
    std::vector<Particle> particles;
    // the file length is checked against the header BEFORE anything is allocated from it
    long long file_bytes = -1;
    if (std::fseek(f, 0, SEEK_END) == 0) file_bytes = (long long)std::ftell(f);
    std::rewind(f);
    bool ok = file_bytes >= (long long)kCheckpointHeaderV1Bytes &&
              std::fread(&h, kCheckpointHeaderV1Bytes, 1, f) == 1 &&
              std::memcmp(h.magic, kMagic, sizeof(kMagic)) == 0 &&
              (h.version == 1 || h.version == 2) && h.num_particles >= 0;
    const size_t header_bytes = h.version == 2 ? sizeof(CheckpointHeader) : kCheckpointHeaderV1Bytes;
    ok = ok && file_bytes == (long long)header_bytes + (long long)h.num_particles * (long long)sizeof(Particle);
    if (ok && h.version == 2)
        ok = std::fread(reinterpret_cast<char*>(&h) + kCheckpointHeaderV1Bytes,
                        sizeof(CheckpointHeader) - kCheckpointHeaderV1Bytes, 1, f) == 1;
    if (ok) {
        particles.resize((size_t)h.num_particles);
        ok = particles.empty() ||
             std::fread(particles.data(), sizeof(Particle), particles.size(), f) == particles.size();
    }
    std::fclose(f);
    if (!ok) throw Error(WC_ERR_INVALID, "loadCheckpoint: " + path + " is not a checkpoint (bad magic, version or length)");
    if (header) *header = h;
    return particles;
}

}  // namespace util
}  // namespace core
