/*
 * oracle/ref_shim_host.cpp -- TEST INFRASTRUCTURE ONLY.
 * extern "C" wrapper (ours) around the REFERENCE's own Core/src/Utils/OdometryProvider.h (rodrigues, computeUpdateSE3: the pose
 * update of the Gauss-Newton loop), included where it lies and compiled unmodified against oracle/eigen_mini
 * (Eigen itself is not installed) by oracle/build_ref_host.sh into oracle/_ref/libref_host.so.
 */
#include "OdometryProvider.h"

extern "C" {
void refh_rodrigues(const double w[3], double R[9])
{
    const Eigen::Matrix<double, 3, 3, Eigen::RowMajor> r = OdometryProvider::rodrigues(Eigen::Vector3d(w[0], w[1], w[2]));
    std::memcpy(R, r.data(), 9 * sizeof(double));
}
/* resultRt: row-major 4x4, in/out; xi = (t, omega); iso16: the Isometry3f the reference hands on, row-major */
void refh_computeUpdateSE3(double resultRt[16], const double xi[6], float iso16[16])
{
    Eigen::Matrix<double, 4, 4, Eigen::RowMajor> Rt;
    std::memcpy(Rt.data(), resultRt, 16 * sizeof(double));
    Eigen::Matrix<double, 6, 1> r;
    for (int k = 0; k < 6; ++k) r(k) = xi[k];
    Eigen::Isometry3f iso;
    iso.setIdentity();
    OdometryProvider::computeUpdateSE3(Rt, r, iso);
    std::memcpy(resultRt, Rt.data(), 16 * sizeof(double));
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) iso16[i * 4 + j] = iso(i, j);
}
}
