// tests/classes_tu.cpp -- drives include/hrbf_classes.hpp exactly as HRBFFusion::processFrame / predict drive the reference's classes
// (Core/src/HRBFFusion.cpp:1006-1052, 1063-1130, 1192-1227, 1244-1260): same calls, same order, same arguments.
// usage: classes_tu W H n_frames frames.bin poses_out.bin [icpWeight so3]     frames.bin = n x { rgb8[H][W][3], depth16[H][W] }
#include <hrbf_classes.hpp>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace hrbf_b200;

static float rodrigues2_norm(const float R[9])      // |log R| as HRBFFusion::rodrigues2 returns it
{
    double c = ((double)R[0] + R[4] + R[8] - 1.0) * 0.5;
    c = c > 1 ? 1 : c < -1 ? -1 : c;
    return (float)std::acos(c);
}

int main(int argc, char** argv)
{
    if (argc < 6) return 2;
    const int W = atoi(argv[1]), H = atoi(argv[2]), n = atoi(argv[3]);
    const float icpWeight = argc > 6 ? (float)atof(argv[6]) : 10.f;
    const bool so3 = argc > 7 ? atoi(argv[7]) != 0 : true;
    const float s = W / 640.f, fx = 528.f * s, fy = 528.f * s, cx = 320.f * s, cy = 240.f * s;
    FILE* f = fopen(argv[4], "rb");
    if (!f) return 3;
    std::vector<unsigned char> rgb((size_t)W * H * 3);
    std::vector<unsigned short> depth((size_t)W * H);

    // the members of HRBFFusion
    FrameTextures textures(W, H, cx, cy, fx, fy);
    RGBDOdometry frameToModel(W, H, cx, cy, fx, fy);
    IndexMap indexMap(W, H, cx, cy, fx, fy);
    GlobalModel globalModel(W, H, cx, cy, fx, fy, 1u << 20);
    FillIn fillIn(W, H);
    Mat4 currPose, lastPose;
    const float maxDepthProcessed = 20.f, confidenceThreshold = 5.f;
    const bool rgbOnly = false, pyramid = true, fastOdom = false, frameToFrameRGB = false, insertSubmap = false, lost = false;
    const int indexSubmap = 0;
    int tick = 1;
    std::vector<float> poses;

    for (int k = 0; k < n; ++k) {
        if (fread(rgb.data(), 1, rgb.size(), f) != rgb.size() || fread(depth.data(), 2, depth.size(), f) != depth.size()) return 4;
        textures.Upload(rgb.data(), depth.data());                                            // :1006-1010
        textures.preprocess();                                                                // :1017-1021
        float weighting = 1.f;
        if (tick == 1) {
            globalModel.initialise(textures[HRBF_FT_VERTEX_RAW], textures[HRBF_FT_NORMAL], textures[HRBF_FT_RGB], textures[HRBF_FT_PRINCIPAL_CURV1],
                                   textures[HRBF_FT_PRINCIPAL_CURV2], textures[HRBF_FT_GRADIENT_MAG], currPose);      // :1043-1049
            frameToModel.initFirstRGB(textures[HRBF_FT_RGBA]);                                // :1052
        } else {
            lastPose = currPose;
            const bool shouldFillIn = !denseEnough(indexMap.vertexTexHRBF());                 // :1069-1070
            frameToModel.initICPModel(shouldFillIn ? &fillIn.vertexTexture : indexMap.vertexTexHRBF(), shouldFillIn ? &fillIn.normalTexture : indexMap.normalTexHRBF(),
                                      maxDepthProcessed, currPose);
            frameToModel.initRGBModel((shouldFillIn || frameToFrameRGB) ? &fillIn.imageTexture : indexMap.imageTexHRBF());
            frameToModel.initCurvatureModel(shouldFillIn ? &fillIn.curvk1Texture : indexMap.curvk1TexHRBF(), shouldFillIn ? &fillIn.curvk2Texture : indexMap.curvk2TexHRBF(), currPose);
            frameToModel.initICP(textures[HRBF_FT_VERTEX_FILTERED], textures[HRBF_FT_NORMAL], maxDepthProcessed);
            frameToModel.initRGB(textures[HRBF_FT_RGBA]);
            frameToModel.initCurvature(textures[HRBF_FT_PRINCIPAL_CURV1], textures[HRBF_FT_PRINCIPAL_CURV2]);
            frameToModel.initICPweight(shouldFillIn ? &fillIn.icpweightTexture : indexMap.icpweightTexHRBF());
            float trans[3] = { currPose(0, 3), currPose(1, 3), currPose(2, 3) };
            float rot[9];
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rot[i * 3 + j] = currPose(i, j);
            frameToModel.getIncrementalTransformation(trans, rot, rgbOnly, icpWeight, pyramid, fastOdom, so3, true, tick - 1);      // :1092-1100
            for (int i = 0; i < 3; ++i) { currPose(i, 3) = trans[i]; for (int j = 0; j < 3; ++j) currPose(i, j) = rot[i * 3 + j]; }
            // weight by velocity (:1112-1123): diff = currPose^-1 * lastPose
            float R[9], t[3];
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) R[i * 3 + j] = currPose(0, i) * lastPose(0, j) + currPose(1, i) * lastPose(1, j) + currPose(2, i) * lastPose(2, j);
                t[i] = currPose(0, i) * (lastPose(0, 3) - currPose(0, 3)) + currPose(1, i) * (lastPose(1, 3) - currPose(1, 3)) + currPose(2, i) * (lastPose(2, 3) - currPose(2, 3));
            }
            weighting = std::fmax(std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]), rodrigues2_norm(R));
            const float largest = 0.01f, minWeight = 0.5f;
            if (weighting > largest) weighting = largest;
            weighting = std::fmax(1.0f - (weighting / largest), minWeight) * 1.0f;
            textures.VertexConfidence(weighting);                                             // :1126
            if (!rgbOnly && !lost) {                                                          // :1192-1227
                indexMap.predictIndices(currPose, tick, tick, globalModel.model(), maxDepthProcessed, insertSubmap, indexSubmap);
                globalModel.fuse(currPose, tick, textures[HRBF_FT_RGB], textures[HRBF_FT_DEPTH_METRIC], textures[HRBF_FT_DEPTH_METRIC_FILTERED], textures[HRBF_FT_PRINCIPAL_CURV1],
                                 textures[HRBF_FT_PRINCIPAL_CURV2], textures[HRBF_FT_CONFIDENCE], indexMap.indexTex(), indexMap.vertConfTex(), indexMap.colorTimeTex(),
                                 indexMap.normalRadTex(), maxDepthProcessed, confidenceThreshold, weighting, insertSubmap, (float)indexSubmap);
                indexMap.predictIndices(currPose, tick, tick, globalModel.model(), maxDepthProcessed, insertSubmap, indexSubmap);
                globalModel.clean(currPose, tick, indexMap.indexTex(), indexMap.vertConfTex(), indexMap.colorTimeTex(), indexMap.normalRadTex(), indexMap.depthTex(),
                                  confidenceThreshold, maxDepthProcessed);
            }
        }
        // predict (:1244-1260)
        indexMap.predictIndices(currPose, tick, tick, globalModel.model(), maxDepthProcessed, insertSubmap, indexSubmap);
        indexMap.predictHRBF(IndexMap::ACTIVE);
        fillIn.run(indexMap, textures, lost);
        ++tick;
        for (int q = 0; q < 16; ++q) poses.push_back(currPose.m[q]);
        printf("frame %d: surfels %u  t = %.6f %.6f %.6f\n", k, globalModel.lastCount(), currPose(0, 3), currPose(1, 3), currPose(2, 3));
    }
    fclose(f);
    FILE* o = fopen(argv[5], "wb");
    fwrite(poses.data(), sizeof(float), poses.size(), o);
    fclose(o);
    return 0;
}
