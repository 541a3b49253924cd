// open3d/Open3D.h — API SHIM, not Open3D (test infrastructure, see Eigen/Dense in this tree).
// Only what the reference's hot-path sources touch.  Arithmetic restated from Open3D's sources:
//   geometry::PointCloud::Transform      p <- (T [p;1])_{0..2} / (T [p;1])_3 ; C <- R C R^T
//   geometry::KDTreeFlann::Search (KNN)  exact k nearest neighbours, k = 30 by default; results in
//                                        ascending (distance^2, index) order (Open3D leaves the order
//                                        of equal distances to nanoflann's traversal)
//   utility::ComputeCovariance           nine raw cumulants -> population covariance
//   utility::hash_eigen                  boost-style hash combine
// Visualiser / mesh / IO classes are empty stand-ins: the wrapper never enables `visualize`.
#ifndef ESKF_REFSHIM_OPEN3D_H_
#define ESKF_REFSHIM_OPEN3D_H_

#include <algorithm>
#include <functional>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include <Eigen/Dense>

namespace open3d
{
namespace geometry
{
class Geometry
{
public:
  virtual ~Geometry() = default;
};

class PointCloud : public Geometry
{
public:
  std::vector<Eigen::Vector3d> points_;
  std::vector<Eigen::Matrix3d> covariances_;

  PointCloud & Transform(const Eigen::Matrix4d & T)
  {
    for (auto & p : points_) {
      // Matrix4d * Vector4d is a packet product in Eigen: per row, multiply-adds in column order
      double h[4];
      for (int i = 0; i < 4; ++i) {h[i] = ((T(i, 0) * p(0) + T(i, 1) * p(1)) + T(i, 2) * p(2)) + T(i, 3) * 1.0;}
      p = Eigen::Vector3d(h[0], h[1], h[2]) / h[3];
    }
    const Eigen::Matrix3d R = T.block<3, 3>(0, 0);
    for (auto & c : covariances_) {c = R * c * R.transpose();}
    return *this;
  }
};

class KDTreeSearchParamKNN
{
public:
  explicit KDTreeSearchParamKNN(int knn = 30) : knn_(knn) {}
  int knn_;
};

class KDTreeFlann
{
public:
  bool SetGeometry(const PointCloud & cloud)
  {
    pts_ = &cloud.points_;
    // a uniform grid over the bounding box (cell ~ cube root of volume per 8 points) only prunes the
    // exact search below; it never changes its result
    n_ = static_cast<int>(pts_->size());
    if (n_ == 0) {return false;}
    for (int d = 0; d < 3; ++d) {lo_[d] = 1e300; hi_[d] = -1e300;}
    for (const auto & p : *pts_) {
      for (int d = 0; d < 3; ++d) {lo_[d] = std::min(lo_[d], p(d)); hi_[d] = std::max(hi_[d], p(d));}
    }
    double vol = 1.0;
    for (int d = 0; d < 3; ++d) {vol *= std::max(hi_[d] - lo_[d], 1e-3);}
    cell_ = std::max(std::cbrt(vol * 8.0 / n_), 1e-3);
    for (int d = 0; d < 3; ++d) {dim_[d] = std::min(512, static_cast<int>((hi_[d] - lo_[d]) / cell_) + 1);}
    start_.assign(static_cast<size_t>(dim_[0]) * dim_[1] * dim_[2] + 1, 0);
    std::vector<int> cell_of(n_);
    for (int i = 0; i < n_; ++i) {cell_of[i] = cellIndex((*pts_)[i]); ++start_[cell_of[i] + 1];}
    for (size_t c = 1; c < start_.size(); ++c) {start_[c] += start_[c - 1];}
    order_.resize(n_);
    std::vector<int> fill(start_.begin(), start_.end() - 1);
    for (int i = 0; i < n_; ++i) {order_[fill[cell_of[i]]++] = i;}
    return true;
  }

  int Search(
    const Eigen::Vector3d & q, const KDTreeSearchParamKNN & param, std::vector<int> & indices,
    std::vector<double> & distance2) const
  {
    const int k = std::min(param.knn_, n_);
    std::vector<std::pair<double, int>> heap;  // max-heap of the best k (distance^2, index)
    heap.reserve(k + 1);
    int c[3];
    coords(q, c);
    // rings of cells around the query's cell; stop once the ring is farther than the k-th distance
    for (int r = 0;; ++r) {
      // a point of ring r differs from the query's cell by r cells along some axis, so it is at
      // least (r - 1) * cell away: once that exceeds the k-th distance the search is complete
      if (static_cast<int>(heap.size()) == k && r > 0) {
        const double reach = (r - 1) * cell_;
        if (reach * reach > heap.front().first) {break;}
      }
      bool any = false;
      for (int x = c[0] - r; x <= c[0] + r; ++x) {
        if (x < 0 || x >= dim_[0]) {continue;}
        for (int y = c[1] - r; y <= c[1] + r; ++y) {
          if (y < 0 || y >= dim_[1]) {continue;}
          for (int z = c[2] - r; z <= c[2] + r; ++z) {
            if (z < 0 || z >= dim_[2]) {continue;}
            if (std::max({std::abs(x - c[0]), std::abs(y - c[1]), std::abs(z - c[2])}) != r) {continue;}
            any = true;
            const size_t ci = (static_cast<size_t>(x) * dim_[1] + y) * dim_[2] + z;
            for (int s = start_[ci]; s < start_[ci + 1]; ++s) {
              const int i = order_[s];
              const Eigen::Vector3d & p = (*pts_)[i];
              const double dx = p(0) - q(0), dy = p(1) - q(1), dz = p(2) - q(2);
              const std::pair<double, int> e((dx * dx + dy * dy) + dz * dz, i);
              if (static_cast<int>(heap.size()) < k) {
                heap.push_back(e);
                std::push_heap(heap.begin(), heap.end());
              } else if (e < heap.front()) {
                std::pop_heap(heap.begin(), heap.end());
                heap.back() = e;
                std::push_heap(heap.begin(), heap.end());
              }
            }
          }
        }
      }
      if (!any && r > std::max({dim_[0], dim_[1], dim_[2]})) {break;}
    }
    std::sort(heap.begin(), heap.end());
    indices.resize(heap.size());
    distance2.resize(heap.size());
    for (size_t j = 0; j < heap.size(); ++j) {indices[j] = heap[j].second; distance2[j] = heap[j].first;}
    return static_cast<int>(heap.size());
  }

private:
  void coords(const Eigen::Vector3d & p, int * c) const
  {
    for (int d = 0; d < 3; ++d) {
      const int v = static_cast<int>(std::floor((p(d) - lo_[d]) / cell_));
      c[d] = std::min(std::max(v, 0), dim_[d] - 1);
    }
  }
  int cellIndex(const Eigen::Vector3d & p) const
  {
    int c[3];
    coords(p, c);
    return (c[0] * dim_[1] + c[1]) * dim_[2] + c[2];
  }
  const std::vector<Eigen::Vector3d> * pts_ = nullptr;
  int n_ = 0;
  double lo_[3], hi_[3], cell_ = 1.0;
  int dim_[3] = {1, 1, 1};
  std::vector<int> start_, order_;
};

class LineSet : public Geometry
{
public:
  std::vector<Eigen::Vector3d> points_;
  std::vector<Eigen::Vector2i> lines_;
  std::vector<Eigen::Vector3d> colors_;
};

class TriangleMesh : public Geometry
{
public:
  static std::shared_ptr<TriangleMesh> CreateCoordinateFrame(double, const Eigen::Vector3d &)
  {
    return std::make_shared<TriangleMesh>();
  }
  TriangleMesh & Rotate(const Eigen::Matrix3d &, const Eigen::Vector3d &) {return *this;}
};
}  // namespace geometry

namespace utility
{
template<typename T>
struct hash_eigen
{
  std::size_t operator()(T const & matrix) const
  {
    size_t seed = 0;
    for (int i = 0; i < static_cast<int>(matrix.size()); i++) {
      auto elem = *(matrix.data() + i);
      seed ^= std::hash<typename T::Scalar>()(elem) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    }
    return seed;
  }
};

template<typename IdxType>
Eigen::Matrix3d ComputeCovariance(const std::vector<Eigen::Vector3d> & points, const std::vector<IdxType> & indices)
{
  if (indices.empty()) {return Eigen::Matrix3d::Identity();}
  Eigen::Matrix3d covariance;
  double c[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (const auto & idx : indices) {
    const Eigen::Vector3d & p = points[idx];
    c[0] += p(0);
    c[1] += p(1);
    c[2] += p(2);
    c[3] += p(0) * p(0);
    c[4] += p(0) * p(1);
    c[5] += p(0) * p(2);
    c[6] += p(1) * p(1);
    c[7] += p(1) * p(2);
    c[8] += p(2) * p(2);
  }
  for (double & v : c) {v /= static_cast<double>(indices.size());}
  covariance(0, 0) = c[3] - c[0] * c[0];
  covariance(1, 1) = c[6] - c[1] * c[1];
  covariance(2, 2) = c[8] - c[2] * c[2];
  covariance(0, 1) = c[4] - c[0] * c[1];
  covariance(1, 0) = covariance(0, 1);
  covariance(0, 2) = c[5] - c[0] * c[2];
  covariance(2, 0) = covariance(0, 2);
  covariance(1, 2) = c[7] - c[1] * c[2];
  covariance(2, 1) = covariance(1, 2);
  return covariance;
}
}  // namespace utility

namespace camera
{
class PinholeCameraParameters
{
public:
  Eigen::Matrix4d extrinsic_ = Eigen::Matrix4d::Identity();
};
class PinholeCameraTrajectory
{
public:
  std::vector<PinholeCameraParameters> parameters_;
};
}  // namespace camera

namespace visualization
{
class RenderOption
{
public:
  void SetPointSize(double) {}
  Eigen::Vector3d background_color_ = Eigen::Vector3d::Zero();
};
class ViewControl
{
public:
  bool ConvertFromPinholeCameraParameters(const camera::PinholeCameraParameters &) {return true;}
};
class Visualizer
{
public:
  bool CreateVisualizerWindow(const std::string &, int, int) {return true;}
  RenderOption & GetRenderOption() {return ro_;}
  ViewControl & GetViewControl() {return vc_;}
  void ClearGeometries() {}
  template<class G> bool AddGeometry(const std::shared_ptr<G> &) {return true;}
  bool PollEvents() {return true;}
  bool HasGeometry() const {return false;}
  void UpdateRender() {}

private:
  RenderOption ro_;
  ViewControl vc_;
};
}  // namespace visualization

namespace io
{
inline bool WritePointCloud(const std::string &, const geometry::PointCloud &) {return false;}
inline bool WritePinholeCameraTrajectory(const std::string &, const camera::PinholeCameraTrajectory &) {return false;}
}  // namespace io
}  // namespace open3d

#endif  // ESKF_REFSHIM_OPEN3D_H_
