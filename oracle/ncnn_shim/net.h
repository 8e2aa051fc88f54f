// TEST INFRASTRUCTURE.  Minimal stand-in for ncnn's "net.h" so that the reference's own deploy post-process
// (/root/reference/sample/ncnn/src/yolo-fastestv2.cpp, compiled where it lies, never copied) builds without ncnn:
// just enough of ncnn::Mat / Net / Extractor for that translation unit.  No inference happens here; the two
// output blobs are whatever the harness (ref_harness.cpp) placed in g_blobs before calling detection().
#ifndef YFV2_ORACLE_NCNN_SHIM_NET_H_
#define YFV2_ORACLE_NCNN_SHIM_NET_H_
#include <assert.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

namespace ncnn {

// dims = 3 blob, dense: `c` channels of h rows of w floats (the sample reads it as c = grid rows, h = grid columns, w = 5A+C)
class Mat {
public:
    int w = 0, h = 0, c = 0;
    std::vector<float> store;
    enum PixelType { PIXEL_BGR = 2 };
    Mat() {}
    Mat(int w_, int h_, int c_) : w(w_), h(h_), c(c_), store((size_t)w_ * h_ * c_) {}
    const float* channel(int q) const { return store.data() + (size_t)q * w * h; }
    float* channel(int q) { return store.data() + (size_t)q * w * h; }
    static Mat from_pixels_resize(const unsigned char*, int, int, int, int, int) { return Mat(); }
    void substract_mean_normalize(const float*, const float*) {}
};

extern std::map<std::string, Mat> g_blobs;      // defined by the harness

class Extractor {
public:
    void set_num_threads(int) {}
    int input(const char*, const Mat&) { return 0; }
    int extract(const char* name, Mat& out) {
        std::map<std::string, Mat>::const_iterator it = g_blobs.find(name);
        if (it == g_blobs.end()) return -1;
        out = it->second;
        return 0;
    }
};

class Net {
public:
    int load_param(const char*) { return 0; }
    int load_model(const char*) { return 0; }
    Extractor create_extractor() { return Extractor(); }
};

}  // namespace ncnn
#endif
