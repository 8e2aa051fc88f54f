// TEST INFRASTRUCTURE.  Stand-in for <opencv2/opencv.hpp>: the reference's ncnn sample only reads cols / rows / data of the
// source image (sample/ncnn/src/yolo-fastestv2.cpp:189-194).
#ifndef YFV2_ORACLE_NCNN_SHIM_OPENCV_HPP_
#define YFV2_ORACLE_NCNN_SHIM_OPENCV_HPP_
namespace cv {
class Mat {
public:
    int cols = 0, rows = 0;
    unsigned char* data = nullptr;
};
}  // namespace cv
#endif
