// Stand-in for the few OpenCV names databases.cpp:103-111 mentions (k-means++ seeding of
// the offline coarse-quantiser training, out of scope). Lets databases.cpp compile so the
// oracle can link substract_vectors_from_unique. TEST INFRASTRUCTURE ONLY.
#ifndef QADC_OPENCV_STUB_HPP
#define QADC_OPENCV_STUB_HPP
#define CV_32F 5
#define CV_32S 4
#define CV_TERMCRIT_ITER 1
namespace cv {
struct Mat { Mat(long, int, int, void*) {} };
struct TermCriteria { TermCriteria(int, int, double) {} };
enum { KMEANS_PP_CENTERS = 2 };
inline double kmeans(const Mat&, int, Mat&, TermCriteria, int, int, Mat&) { return 0.0; }
}  // namespace cv
#endif
