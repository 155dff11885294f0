// map_stitcher.hpp -- header-only C++ shim with MapStitcher's call surface (include/map_stitcher.h:24-41) on top of the C ABI
// (nis_stitcher_*, include/nislam.h).
//
//   reference                                                     this shim (nislam::MapStitcherT<Vector3d, ArrayXXi>)
//   ------------------------------------------------------------  -------------------------------------------------------------
//   MapStitcher(MapStitcherConfig config, CameraPtr camera)        MapStitcherT(cell_size, stitch_map, camera model, H, W, cell window)
//   void InsertFrame(FramePtr frame, cv::Mat& image)               InsertFrame(frame, image_u8)   frame: anything with GetPose(Vector3d&)
//   void RecomputeOccupancy()                                      RecomputeOccupancy(frames)     frames in insertion order
//   OccupancyData& GetOccupancyData()                              GetCell(cell_x, cell_y, data, weight) -> bool   one Cell at a time
//
// Differences (documented in INTEGRATION.md): the cells live in a dense window chosen at construction; RecomputeOccupancy replays the
// frames in insertion order (the reference iterates an unordered_map keyed by FramePtr); the frames are passed in because the GPU
// side keeps the scaled images, not the FramePtr keys.  `_to_stitch == false` makes every call a no-op like the reference (:15).
#pragma once
#include <stdexcept>
#include <vector>

#include "../../include/nislam.h"

namespace nislam {

template <class Vector3d, class ArrayXXi>
class MapStitcherT {
 public:
  MapStitcherT(int cell_size, bool stitch_map, const nis_camera_model& camera, int image_height, int image_width, int cell_x0, int cell_y0,
               int cells_x, int cells_y, int device = 0)
      : cell_size_(cell_size), to_stitch_(stitch_map), cam_(camera) {
    if (!to_stitch_) return;
    check(nis_stitcher_create(device, image_height, image_width, cell_size, cell_x0, cell_y0, cells_x, cells_y, &st_));
  }
  ~MapStitcherT() { nis_stitcher_destroy(st_); }
  MapStitcherT(const MapStitcherT&) = delete;
  MapStitcherT& operator=(const MapStitcherT&) = delete;

  // map_stitcher.cc:14-22; image = the undistorted u8 frame, row-major H x W (cv::Mat::data of a continuous CV_8UC1 Mat)
  template <class FramePtr>
  void InsertFrame(const FramePtr& frame, const uint8_t* image_u8) {
    if (!to_stitch_) return;
    Vector3d p;
    frame->GetPose(p);
    const double pose[3] = {p[0], p[1], p[2]};
    check(nis_stitcher_insert(st_, image_u8, pose, &cam_, nullptr));
  }
  // map_stitcher.cc:135-145 with the frames' current (optimised) poses
  template <class FramePtr>
  void RecomputeOccupancy(const std::vector<FramePtr>& frames_in_insertion_order) {
    if (!to_stitch_) return;
    if ((int)frames_in_insertion_order.size() != nis_stitcher_frames(st_)) throw std::invalid_argument("RecomputeOccupancy: frame count mismatch");
    std::vector<double> poses;
    for (const FramePtr& f : frames_in_insertion_order) {
      Vector3d p;
      f->GetPose(p);
      poses.push_back(p[0]); poses.push_back(p[1]); poses.push_back(p[2]);
    }
    check(nis_stitcher_recompute(st_, poses.data(), &cam_));
  }
  // one Cell of GetOccupancyData(): data(y, x) / weight(y, x) like the reference's Eigen::ArrayXXi (resized to cell_size x cell_size)
  bool GetCell(int cell_x, int cell_y, ArrayXXi& data, ArrayXXi& weight) {
    if (!to_stitch_) return false;
    std::vector<int32_t> d((size_t)cell_size_ * cell_size_), w(d.size());
    int present = 0;
    check(nis_stitcher_cell(st_, cell_x, cell_y, d.data(), w.data(), &present));
    if (!present) return false;
    data.resize(cell_size_, cell_size_);
    weight.resize(cell_size_, cell_size_);
    for (int y = 0; y < cell_size_; ++y)
      for (int x = 0; x < cell_size_; ++x) {
        data(y, x) = d[(size_t)y * cell_size_ + x];
        weight(y, x) = w[(size_t)y * cell_size_ + x];
      }
    return true;
  }
  long long DroppedPixels() {
    long long n = 0;
    if (to_stitch_) check(nis_stitcher_dropped(st_, &n));
    return n;
  }

 private:
  static void check(int st) {
    if (st != NIS_OK) throw std::runtime_error(std::string("libnislam: ") + nis_strerror(st));
  }
  int cell_size_;
  bool to_stitch_;
  nis_camera_model cam_;
  nis_stitcher* st_ = nullptr;
};

}  // namespace nislam
