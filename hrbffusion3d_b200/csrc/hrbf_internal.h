// hrbf_internal.h -- library-internal declarations shared by the translation units of libhrbf_b200.so
// (not part of the C ABI; the public surface is include/hrbf_b200.h).
#pragma once
#include "common.cuh"
#include <map>
#include <utility>

namespace hrbf {
struct ReduceWork;
struct TrackState;
enum { M_VG = 0, M_NG, M_K1G, M_K2G, M_VC, M_NC, M_K1C, M_K2C, M_W, M_COUNT };
struct So3Pre;
// Everything the tracker needs that depends on the CURRENT camera frame alone: the outputs of initICP / initRGB / initCurvature
// (RGBDOdometry.cpp:183-247, 689-794), the Sobel images and candidate mask of computeRgbResidual, and the SO3 pre-alignment
// (RGBDOdometry.cpp:827-914: it compares two CAMERA images).  Two banks: the frame pipeline builds bank (t+1) & 1 on its staging
// stream while frame t is tracked from bank t & 1 (odom_stage_current_dev); the stand-alone API only ever uses bank 0.
struct CurrBank {
    float* maps[4][HRBF_NUM_PYRS] = {};          // M_VC, M_NC, M_K1C, M_K2C
    float4* pk[2][HRBF_NUM_PYRS] = {};           // packed current records
    unsigned char* nextImage[HRBF_NUM_PYRS] = {}; float* nextDepth[HRBF_NUM_PYRS] = {};
    short* dIdx[HRBF_NUM_PYRS] = {}; short* dIdy[HRBF_NUM_PYRS] = {};
    unsigned char* cand[HRBF_NUM_PYRS] = {};
    So3Pre* so3 = nullptr;                       // device: result of the staged SO3 pre-alignment
    void* tmaps = nullptr;                       // host tensor maps over this bank's records
    bool cand_ready = false, so3_ready = false;  // staged for the frame this bank holds
    bool image_ready = false;                    // nextImage[] was built by so3_image_kernel (odom_stage_so3_dev): the pyramid job only adds the depth
};
}

struct hrbf_odometry {
    int width = 0, height = 0;
    hrbf_camera intr{};
    float distThres = 0.1f, angleThres = 0.f;
    float sobelScale = 0.125f, maxDepthDeltaRGB = 0.07f, maxDepthRGB = 6.0f;
    float minGrad[HRBF_NUM_PYRS] = { 5, 3, 1 };
    float curvThr = 300.f;
    int useSearch = 0, searchRadius = 2, rgbGradWeight = 0;

    char* slab = nullptr;        // one allocation for everything below
    float* maps[hrbf::M_COUNT][HRBF_NUM_PYRS] = {};
    float* vdepth_tmp = nullptr; // verticesToDepth of the last init_icp* texture
    float* depth_tmp[HRBF_NUM_PYRS] = {};
    float* lastDepth[HRBF_NUM_PYRS] = {}; float* nextDepth[HRBF_NUM_PYRS] = {};
    unsigned char* lastImage[HRBF_NUM_PYRS] = {}; unsigned char* nextImage[HRBF_NUM_PYRS] = {}; unsigned char* lastNextImage[HRBF_NUM_PYRS] = {};
    short* dIdx[HRBF_NUM_PYRS] = {}; short* dIdy[HRBF_NUM_PYRS] = {};
    float* cloud[HRBF_NUM_PYRS] = {};
    float4* pk[4][HRBF_NUM_PYRS] = {};           // packed ICP operands: [0] curr pk0, [1] curr pk1, [2] model pk0, [3] model pk1
    bool pack_dirty_curr = false, pack_dirty_model = false;   // SoA written by a builder that does not pack (GPUTest path)
    unsigned char* cand[HRBF_NUM_PYRS] = {};     // persistent tracker: pose-independent candidate mask of computeRgbResidual
    void* tmaps_host = nullptr;                  // host: CUtensorMap[2 geometries][3 levels][5 arrays] for the TMA-staged ICP tiles (icp_tile.cuh), copied into kernel parameters; null = unavailable
    bool tile_resident = false;                  // persistent tracker: keep each level's ICP tile in shared memory (hrbf_odometry_set_tracker_tiles; measured slower, off)
    int track_threads = 512;                     // threads per CTA of the persistent tracker (256 | 384 | 512), hrbf_odometry_set_tracker_threads
    hrbf_dataterm* corresImg[HRBF_NUM_PYRS] = {};
    hrbf::ReduceWork* work = nullptr;
    unsigned long long *tp_ll_f = nullptr, *tp_ll_i = nullptr;   // persistent tracker: tagged-word exchange buffers
    bool tp_no_pdl = false;                                      // the driver refused cooperative + programmatic serialization
    unsigned int tp_epoch = 0;                                   // launch counter feeding the tags
    int num_sms = 0;
    long long* tp_dbg = nullptr;
    bool use_graph = false;          // true: one kernel per reduction replayed as a CUDA graph (first-generation path)
    float* pose_scratch = nullptr;   // device: [0..11] model pose (R,t), [12..23] track in, [24..35] track out
    // pinned host mirrors
    float* h_pose = nullptr;         // [0..11] in, [12..23] out
    hrbf::TrackState* h_state = nullptr;
    float* h_model_pose = nullptr;   // staging ring for init_*_model poses
    hrbf::SlotRing model_pose_ring;   // guards h_model_pose
    int so3_parity = 0;
    hrbf::CurrBank bank[2];          // bank[0] = the buffers of the slab; bank[1] allocated by odom_enable_banks
    char* bank1_slab = nullptr;
    int cur_bank = 0;
    bool banked = false;

    cudaStream_t cap_stream = nullptr;
    std::map<uint64_t, std::pair<cudaGraphExec_t, int>> graphs;   // key -> (exec, kernel nodes)

    int rows(int l) const { return height >> l; }
    int cols(int l) const { return width >> l; }
};


struct hrbf_indexmap {
    int width = 0, height = 0;
    float cx = 0, cy = 0, fx = 0, fy = 0;
    char* slab = nullptr;
    unsigned long long* keys = nullptr;
    void* tex[HRBF_TEX_COUNT] = {};
    float* active_kf = nullptr;          // device float[HRBF_ACTIVE_KEYFRAME_DIMENSION]
    float* inv_pose = nullptr;           // device: ring of 8 x (Ri[9], ti[3])
    unsigned int* count_slot = nullptr;  // device ring of 8 counts (host-count API)
    float* h_stage = nullptr;            // pinned ring: 8 x 16 floats
    float* h_kf = nullptr;               // pinned keyframe mask
    hrbf::SlotRing ring;                 // guards h_stage / inv_pose / count_slot
    unsigned long long* row_lut = nullptr;      // device: make_pred_row_lut()
    unsigned int* dense_count_next = nullptr;   // frame pipeline: where the next ACTIVE prediction counts its dense-enough samples (or null)
};


namespace hrbf {
// out_mask: 1 colorTime | 2 normRad | 4 curvature maps (index and vertConf are always written); 7 = everything (IndexMap::predictIndices)
int indexmap_splat(hrbf_indexmap* m, const float* inv_pose_dev, const float* surfels, const unsigned int* count_dev, unsigned int bound,
                   float depthCutoff, cudaStream_t s, int out_mask = 7);
// device-pose / device-select forms of the RGBDOdometry init* calls (odometry.cu), used by fusion.cu:
// when `sel` is non-null and *sel != 0 the `_alt` textures are read instead (shouldFillIn, HRBFFusion.cpp:1069-1086)
int odom_init_icp_model_dev(hrbf_odometry* o, const float* v, const float* n, const float* v_alt, const float* n_alt, const int* sel,
                            const float* pose_dev, cudaStream_t s);
int odom_init_curvature_model_dev(hrbf_odometry* o, const float* k1, const float* k2, const float* k1_alt, const float* k2_alt, const int* sel,
                                  const float* pose_dev, cudaStream_t s);
int odom_init_icp_weight_dev(hrbf_odometry* o, const float* w, const float* w_alt, const int* sel, cudaStream_t s);
int odom_init_rgb_model_dev(hrbf_odometry* o, const unsigned char* rgba, const unsigned char* rgba_alt, const int* sel, cudaStream_t s);
// all seven init* calls of one frame (initICPModel, initRGBModel, initCurvatureModel, initICP, initRGB, initCurvature,
// initICPweight; HRBFFusion.cpp:1073-1099) as ONE launch.  Textures as in the single calls; `_alt` = fill-in alternates.
struct OdomPrepInputs {
    const float *vm, *nm, *vm_alt, *nm_alt;          // model vertex / normal (RGBA32F)
    const float *k1m, *k2m, *k1m_alt, *k2m_alt;      // model curvature
    const float *w, *w_alt;                          // model icp weight (R32F)
    const unsigned char *rgba_m, *rgba_m_alt;        // model image (RGBA8)
    const float *vc, *nc, *k1c, *k2c;                // current frame
    const unsigned char* rgba_c;
    const int* sel;                                  // device fill-in decision (or null, then dense_count decides)
    const unsigned int* dense_count;                 // samples with a predicted surface on the 1/20 grid, counted by the prediction kernel
    unsigned int* dense_count_reset;                 // the counter the NEXT prediction will use: zeroed here
    float dense_thresh;                              // globalDenseEnoughThresh
    const float* pose_dev;                           // device model pose R[9], t[3]
    const unsigned char* rgb8_c;                     // current image as RGB8 (used instead of rgba_c when non-null)
};
int odom_prep_all_dev(hrbf_odometry* o, const OdomPrepInputs& in, cudaStream_t s);
// Frame pipeline: two CurrBanks (see above).  odom_stage_current_dev builds bank `b` from a preprocessed frame's textures -- the
// current-frame jobs of prep_all, Sobel + candidates -- on any stream; odom_select_bank makes a bank the one the tracker and the
// init* calls use.
int odom_enable_banks(hrbf_odometry* o);
void odom_select_bank(hrbf_odometry* o, int b);
int odom_stage_current_dev(hrbf_odometry* o, int b, const OdomPrepInputs& in, cudaStream_t s);
// The SO3 pre-alignment of the frame in bank b against the previous camera frame (bank b ^ 1): needs the uploaded RGB8 only, so
// it can run beside the preprocessing (its <= 10 dependent reductions are latency, not work).
// image_done (or null) is recorded once the bank's intensity pyramid is written (odom_stage_current_dev reads it).
int odom_stage_so3_dev(hrbf_odometry* o, int b, const unsigned char* rgb8, bool so3, bool has_previous, cudaEvent_t image_done, cudaStream_t s);
struct OdomFrameEpilogue { float* last_pose_out; float* inv_pose_out; float* weighting_out; float weight_multiplier; float* traj_out; };
int odom_track_frame_dev(hrbf_odometry* o, float* pose_inout, const OdomFrameEpilogue& ep, bool rgbOnly, float icpWeight, bool pyramid,
                         bool fastOdom, bool so3, bool use_weight, cudaStream_t s);
}
