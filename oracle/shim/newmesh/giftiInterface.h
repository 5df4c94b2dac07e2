// TEST INFRASTRUCTURE ONLY — inert stand-ins for FSL's NEWMESH GIFTI wrapper so that the
// reference's mesh.cpp compiles. File I/O is outside the hot path; the
// oracle driver feeds meshes through the in-memory API and never touches these.
#ifndef ORACLE_SHIM_GIFTI_H
#define ORACLE_SHIM_GIFTI_H
#include <fstream>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#define GIFTI_IND_ORD_ROW_MAJOR 1
#define GIFTI_ENCODING_B64GZ 3

namespace NiftiIO {
enum { NIFTI_INTENT_NONE = 0, NIFTI_INTENT_POINTSET = 1008, NIFTI_INTENT_TRIANGLE = 1009 };
enum { NIFTI_TYPE_INT32 = 8, NIFTI_TYPE_FLOAT32 = 16 };
}

namespace NEWMESH {

struct GIFTImeta { std::string name, value; };
struct GIFTIlabel { std::string name; float RGBA[4]; };

struct GIFTIcoordinateSystem {
    std::string dataSpace, transformSpace;
    std::vector<double> transform;
    GIFTIcoordinateSystem() = default;
    GIFTIcoordinateSystem(const std::string& d, const std::string& t, const std::vector<double>& x)
        : dataSpace(d), transformSpace(t), transform(x) {}
};

class GIFTIfield {
public:
    GIFTIfield() = default;
    GIFTIfield(int, int, int, const int*, const void*, int,
               const std::vector<GIFTIcoordinateSystem>& = std::vector<GIFTIcoordinateSystem>()) {}
    int getDim(int) const { return 0; }
    std::vector<float> fVector(int) const { return {}; }
    std::vector<int> iVector(int) const { return {}; }
    float fScalar(int) const { return 0.f; }
    std::vector<GIFTIcoordinateSystem> getCoordSystems() const { return {}; }
};

class GIFTIwrapper {
public:
    std::vector<GIFTImeta> metaData, extraAttributes;
    std::map<int, GIFTIlabel> GIFTIlabels;
    std::vector<GIFTIfield> allFields;
    void readGIFTI(const std::string&) { throw std::runtime_error("shim GIFTI: file I/O is not provided"); }
    // The groupwise driver never calls set_output_format (src/newmsm.cpp:14-28), so its outputs always go through save_gifti.
    // The stand-in writes a placeholder so that the program runs to completion; results are compared through the trace hook.
    void writeGIFTI(const std::string& f, int) {
        std::ofstream o(f.c_str());
        o << "GIFTI output is not provided by the FSL stand-in (oracle/shim); see the ASCII outputs / MSMGPU_TRACE\n";
    }
    std::vector<GIFTIfield> returnSurfaceFields() const { return {}; }
    std::vector<GIFTIfield> returnNonSurfaceFields() const { return {}; }
};

} // namespace NEWMESH
#endif
