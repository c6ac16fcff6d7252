// TEST-ONLY stand-in for <pcl/point_types.h> (PCL and Eigen are not installed in this image).
// The unmodified reference (ikd_Tree.h:11,62 and ikd_Tree.cpp:1447-1449) needs only
// Eigen::aligned_allocator and three pcl point structs with public float x,y,z.
// This file is oracle/test infrastructure; it is never included by the product.
#pragma once
#include <vector>
#include <memory>
#include <cstring>
#include <cstdint>
#include <cmath>
namespace Eigen { template <class T> using aligned_allocator = std::allocator<T>; }
namespace pcl {
struct alignas(16) PointXYZ { float x = 0, y = 0, z = 0, pad_ = 1.0f; };
struct alignas(16) PointXYZI { float x = 0, y = 0, z = 0, pad_ = 1.0f; float intensity = 0; float pad2_[3] = {0, 0, 0}; };
struct alignas(16) PointXYZINormal {
    float x = 0, y = 0, z = 0, pad_ = 1.0f;
    float normal_x = 0, normal_y = 0, normal_z = 0, pad3_ = 0;
    float intensity = 0, curvature = 0, pad4_[2] = {0, 0};
};
}  // namespace pcl
