// TEST INFRASTRUCTURE (oracle/sdf_ref): stand-in for the three Sophus calls of layers/sdf_matching_loss_kernel.cu
// (the vendored Sophus/sophus headers need all of Eigen, which is absent here).  Sophus only forwards to Eigen's
// quaternion on this path:
//   SE3(Matrix4)      -> so3_(T.topLeftCorner<3,3>()) -> unit_quaternion_(R)   Sophus/sophus/se3.hpp:387-389, so3.hpp:392
//   so3().matrix()    -> unit_quaternion_.toRotationMatrix()                    so3.hpp:264-266
//   SE3 * point       -> so3() * p + translation(), so3 * p = q._transformVector(p)   se3.hpp:249-251, so3.hpp:298-300
// (the orthogonality SOPHUS_ENSUREs only printf in device code, common.hpp:119-122)
#pragma once
#include <Eigen/Core>

namespace Sophus {

template <typename T> struct SO3 {
    Eigen::Quaternion<T> unit_quaternion_;
    OMG_SHIM_HD SO3() {}
    template <int Opt> OMG_SHIM_HD explicit SO3(const Eigen::Matrix<T, 3, 3, Opt> &R) : unit_quaternion_(R) {}
    OMG_SHIM_HD Eigen::Matrix<T, 3, 3, 0> matrix() const { return unit_quaternion_.toRotationMatrix(); }
    template <int Opt> OMG_SHIM_HD Eigen::Matrix<T, 3, 1, Opt> operator*(const Eigen::Matrix<T, 3, 1, Opt> &p) const {
        return unit_quaternion_._transformVector(p);
    }
};

template <typename T> struct SE3 {
    SO3<T> so3_;
    T translation_[3];
    template <int Opt> OMG_SHIM_HD explicit SE3(const Eigen::Matrix<T, 4, 4, Opt> &M) {
        Eigen::Matrix<T, 3, 3, 0> R;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R(i, j) = M(i, j);
        so3_ = SO3<T>(R);
        for (int i = 0; i < 3; ++i) translation_[i] = M(i, 3);
    }
    OMG_SHIM_HD const SO3<T> &so3() const { return so3_; }
    template <int Opt> OMG_SHIM_HD Eigen::Matrix<T, 3, 1, Opt> operator*(const Eigen::Matrix<T, 3, 1, Opt> &p) const {
        Eigen::Matrix<T, 3, 1, Opt> r = so3_ * p;
        for (int i = 0; i < 3; ++i) r.d[i] = r.d[i] + translation_[i];
        return r;
    }
};

}  // namespace Sophus
