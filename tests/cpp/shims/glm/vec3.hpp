// Test stand-in for glm's vec3 (glm is not installed in this image): just enough for the glm overloads of
// include/chad/tsdf.hpp to be compiled and exercised by tests/cpp/facade_overloads.cpp. Not part of the product.
#pragma once
namespace glm {
    struct vec3 {
        float x, y, z;
        vec3() : x(0), y(0), z(0) {}
        vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    };
}
