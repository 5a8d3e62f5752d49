// core/Scene.h -- the object registry that drives the per-frame update.
//
// Same surface as the reference's scene manager (src/core/Scene.h:23-49): objects are
// registered once under their unique name, update() and reset() visit every object in
// registration order (Scene.cpp:52-56, :64-68), draw() only the visible ones.  With it the
// caller of the hot path reads exactly like WaterCubeApp (WaterCubeApp.cpp:59-66, :98):
//
//     scene->addObject(fluid);      // setup
//     scene->update(frame_time);    // every frame -> Fluid::update -> wc_step
//
// Header-only; no Cinder types.  One registry of (object, visible) records plus a name index
// replaces the reference's three parallel containers (its display set is ordered by pointer
// value, so its draw order is arbitrary; here it is registration order).
#pragma once

#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "./BaseObject.h"

namespace core {

typedef std::shared_ptr<class Scene> SceneRef;

class Scene {
public:
    static SceneRef create() { return std::make_shared<Scene>(); }

    int numObjects() const { return (int)entries_.size(); }
    bool exists(const std::string& name) const { return index_.count(name) != 0; }

    // false for a null object or a name that is already taken (Scene.cpp:20-38)
    bool addObject(BaseObjectRef object, bool visible = true) {
        if (!object || exists(object->name())) return false;
        index_[object->name()] = entries_.size();
        entries_.push_back(Entry{object, visible});
        return true;
    }

    BaseObjectRef getObject(const std::string& name) const {
        auto it = index_.find(name);
        return it == index_.end() ? BaseObjectRef() : entries_[it->second].object;
    }
    BaseObjectRef getObjectFromIndex(unsigned int index) const {
        return index < entries_.size() ? entries_[index].object : BaseObjectRef();
    }

    void update(double time) {
        for (const Entry& e : entries_) e.object->update(time);
    }
    void draw() {
        for (const Entry& e : entries_)
            if (e.visible) e.object->draw();
    }
    void reset() {
        for (const Entry& e : entries_) e.object->reset();
    }
    void clear() {
        entries_.clear();
        index_.clear();
    }

private:
    struct Entry {
        BaseObjectRef object;
        bool visible;
    };
    std::vector<Entry> entries_;
    std::unordered_map<std::string, size_t> index_;
};

}  // namespace core
