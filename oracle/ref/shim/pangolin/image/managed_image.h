// Shim for include/visnav/keypoints.h:40 (Pangolin is not in this image).  The reference's keypoint code only
// reads pixels through operator()(x, y) and tests InBounds(x, y, border); this is a non-owning view with those
// two members and the public fields the reference touches (ptr, w, h, pitch in BYTES, as Pangolin has it).
#pragma once
#include <cstddef>
#include <cstdint>
namespace pangolin {
template <class T>
struct ManagedImage {
  T* ptr = nullptr;
  size_t pitch = 0, w = 0, h = 0;
  ManagedImage() = default;
  ManagedImage(T* p, size_t w_, size_t h_, size_t pitch_bytes) : ptr(p), pitch(pitch_bytes), w(w_), h(h_) {}
  const T& operator()(size_t x, size_t y) const {
    return *reinterpret_cast<const T*>(reinterpret_cast<const uint8_t*>(ptr) + y * pitch + x);
  }
  bool InBounds(float x, float y, float border) const {
    return border <= x && x < (float(w) - border) && border <= y && y < (float(h) - border);
  }
};
}  // namespace pangolin
