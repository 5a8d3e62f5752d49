// core/BaseObject.h -- the virtual update/draw/reset hook the scene graph calls
// (reference: src/core/BaseObject.h:24-26, reached from Scene::update, Scene.cpp:52-56).
#pragma once

#include <memory>
#include <string>

namespace core {

typedef std::shared_ptr<class BaseObject> BaseObjectRef;

class BaseObject {
public:
    explicit BaseObject(const std::string& name) : name_(name) {}
    virtual ~BaseObject() {}

    std::string name() const { return name_; }

    virtual void update(double /*time*/) {}
    virtual void draw() {}
    virtual void reset() {}

protected:
    std::string name_;
};

}  // namespace core
