// TEST INFRASTRUCTURE ONLY — stand-in for FSL MISCMATHS (not installed, not in /root/reference).
// Only what the reference's resampler/meshreg hot-path sources name is provided.
#ifndef ORACLE_SHIM_MISCMATHS_H
#define ORACLE_SHIM_MISCMATHS_H
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include "armawrap/newmat.h"

namespace MISCMATHS {

// Whitespace-separated numeric text file -> dense matrix (rows = lines).
inline NEWMAT::Matrix read_ascii_matrix(const std::string& filename) {
    std::ifstream f(filename.c_str());
    std::vector<std::vector<double>> rows;
    std::string line;
    while (std::getline(f, line)) {
        std::istringstream ss(line);
        std::vector<double> r; double v;
        while (ss >> v) r.push_back(v);
        if (!r.empty()) rows.push_back(r);
    }
    NEWMAT::Matrix m((int)rows.size(), rows.empty() ? 0 : (int)rows[0].size());
    for (size_t i = 0; i < rows.size(); ++i)
        for (size_t j = 0; j < rows[i].size() && j < rows[0].size(); ++j) m((int)i + 1, (int)j + 1) = rows[i][j];
    return m;
}

template <class T> inline T Max(const T& a, const T& b) { return a > b ? a : b; }
template <class T> inline T Min(const T& a, const T& b) { return a < b ? a : b; }

} // namespace MISCMATHS
#endif
