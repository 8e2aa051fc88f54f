// TEST INFRASTRUCTURE.  C entry point around the UNMODIFIED reference class yoloFastestv2 (sample/ncnn/src/yolo-fastestv2.cpp,
// compiled from /root/reference by oracle/Makefile `ref`): feeds two [h][w][5A+C] blobs to its detection() (predHandle :134-183
// + nmsHandle :78-110) through the stub extractor and returns the TargetBox list.  Used to pin oracle/ncnn_post.c and to
// generate tests/golden/ncnn_post.npz; never shipped, never measured.
#include <stdio.h>

#include "yolo-fastestv2.h"

namespace ncnn {
std::map<std::string, Mat> g_blobs;
}

static yoloFastestv2* g_det = nullptr;

// Optional: the sample hard-codes 3 anchors, 80 classes, 352x352, NMS 0.25 and the COCO anchors in its constructor (:6-38); this
// setter overwrites those (private) members -- the translation units are built with -Dprivate=public -- so the same reference code
// can be driven at other shapes.  anchors: 2 * A * 2 floats.
extern "C" void ncnn_ref_configure(int A, int C, int in_w, int in_h, float nms_thresh, const float* anchors) {
    if (!g_det) {
        FILE* keep = stdout;
        stdout = fopen("/dev/null", "w");
        g_det = new yoloFastestv2();
        fclose(stdout);
        stdout = keep;
    }
    g_det->numAnchor = A; g_det->numCategory = C; g_det->inputWidth = in_w; g_det->inputHeight = in_h; g_det->nmsThresh = nms_thresh;
    g_det->anchor.assign(anchors, anchors + 4 * A);
}

extern "C" int ncnn_ref_detect(const float* out2, int h2, int w2, const float* out3, int h3, int w3, int ch, int src_cols, int src_rows,
                               float thresh, int max_out, int* boxes /*[max_out][4]*/, float* scores, int* cates) {
    if (!g_det) {
        FILE* keep = stdout;                      // the constructor prints a banner
        stdout = fopen("/dev/null", "w");
        g_det = new yoloFastestv2();
        fclose(stdout);
        stdout = keep;
    }
    ncnn::Mat a(ch, w2, h2), b(ch, w3, h3);
    memcpy(a.store.data(), out2, sizeof(float) * (size_t)ch * w2 * h2);
    memcpy(b.store.data(), out3, sizeof(float) * (size_t)ch * w3 * h3);
    ncnn::g_blobs["794"] = a;                     // outputName1 / outputName2 of the sample (:27-28)
    ncnn::g_blobs["796"] = b;
    cv::Mat src;
    src.cols = src_cols; src.rows = src_rows;
    std::vector<TargetBox> dst;
    g_det->detection(src, dst, thresh);
    int n = (int)dst.size();
    for (int i = 0; i < n && i < max_out; ++i) {
        boxes[4 * i] = dst[i].x1; boxes[4 * i + 1] = dst[i].y1; boxes[4 * i + 2] = dst[i].x2; boxes[4 * i + 3] = dst[i].y2;
        scores[i] = dst[i].score; cates[i] = dst[i].cate;
    }
    return n;
}
