// Projtable.h -- projection-matrix files and trajectories as the reference's tools exchange them
// (HeaderOnly/Utils/Projtable.hxx:138-220, text conventions of LibProjectiveGeometry/EigenToStr.hxx:135-151):
// ".ompl" = one matrix per line, "[p00 p01 p02 p03; p10 p11 p12 p13; p20 p21 p22 p23] " with 12 significant digits;
// lines starting with '#' are comments, "#> key="value" ..." lines carry meta attributes (pixel spacing, detector size).
// Files written here load in the reference's tools and vice versa.
#ifndef ECC_FACADE_PROJTABLE_H
#define ECC_FACADE_PROJTABLE_H

#include <fstream>
#include <iomanip>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../ecc_b200.h"
#include "Compat.h"

namespace ProjTable {

/// "[a b c d; e f g h; i j k l] " -- std::setprecision(12), default float format (EigenToStr.hxx:135-142)
inline std::string toString(const Geometry::ProjectionMatrix& P)
{
    std::ostringstream s;
    s << std::setprecision(12) << "[";
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 4; c++) s << P(r, c) << (c < 3 ? " " : "");
        s << (r < 2 ? "; " : "] ");
    }
    return s.str();
}

/// Brackets, semicolons, commas and tabs count as blanks; anything but twelve numbers gives [I|0] (EigenToStr.hxx:135-151).
inline Geometry::ProjectionMatrix stringToProjectionMatrix(const std::string& in)
{
    Geometry::ProjectionMatrix P = Geometry::ProjectionMatrix::Zero();
    P(0, 0) = P(1, 1) = P(2, 2) = 1;
    std::string s = in;
    for (size_t i = 0; i < s.size(); i++)
        if (s[i] == '\n' || s[i] == '\t' || s[i] == '[' || s[i] == ']' || s[i] == ';' || s[i] == ',') s[i] = ' ';
    std::istringstream str(s);
    std::vector<double> raw;
    double v;
    while (str >> v) raw.push_back(v);
    if (raw.size() == 12)
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 4; c++) P(r, c) = raw[r * 4 + c];
    return P;
}

/// key="value" key2="value2" ... (Projtable.hxx:190-196)
inline std::string toMetaAttrib(const std::map<std::string, std::string>& meta)
{
    std::ostringstream s;
    for (std::map<std::string, std::string>::const_iterator it = meta.begin(); it != meta.end(); ++it)
        s << it->first << "=\"" << it->second << "\" ";
    return s.str();
}

inline void parseMetaAttrib(const std::string& line, std::map<std::string, std::string>& meta)
{
    size_t pos = 0;
    for (;;) {
        const size_t eq = line.find("=\"", pos);
        if (eq == std::string::npos) return;
        size_t k0 = line.find_last_of(" \t", eq);
        k0 = (k0 == std::string::npos || k0 < pos) ? pos : k0 + 1;
        const size_t end = line.find('"', eq + 2);
        if (end == std::string::npos) return;
        meta[line.substr(k0, eq - k0)] = line.substr(eq + 2, end - eq - 2);
        pos = end + 1;
    }
}

/// One matrix per line; '#' comments, "#> " meta lines, the first plain comment is kept as meta["comment"]
/// (Projtable.hxx:168-188).
inline std::vector<Geometry::ProjectionMatrix> loadProjectionsOneMatrixPerLine(const std::string& file,
                                                                               std::map<std::string, std::string>* meta = 0x0)
{
    std::vector<Geometry::ProjectionMatrix> ret;
    std::ifstream pt(file.c_str());
    std::string line;
    while (pt && std::getline(pt, line)) {
        if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
        if (line.empty()) continue;
        if (line[0] == '#') {
            if (meta && line.size() > 1 && line[1] == '>') parseMetaAttrib(line.size() > 3 ? line.substr(3) : std::string(), *meta);
            else if (meta && meta->find("comment") == meta->end()) (*meta)["comment"] = line.substr(1);
        } else
            ret.push_back(stringToProjectionMatrix(line));
    }
    return ret;
}

/// Optional first comment line, optional "#> spacing=... detector_size_px=..." line, then the matrices (Projtable.hxx:198-220).
inline bool saveProjectionsOneMatrixPerLine(const std::vector<Geometry::ProjectionMatrix>& Ps, const std::string& path,
                                            const std::string& first_line_comment = "", double spacing = 0.0, int detector_w = 0,
                                            int detector_h = 0)
{
    std::ofstream file(path.c_str());
    if (!file) return false;
    if (!first_line_comment.empty()) file << "#" << first_line_comment << std::endl;
    if (spacing != 0.0) {
        file << "#> " << "spacing=\"" << spacing << "\"";
        if (detector_w != 0 || detector_h != 0) file << " detector_size_px=\"" << detector_w << " " << detector_h << "\"";
        file << std::endl;
    }
    for (size_t i = 0; i < Ps.size(); i++) file << toString(Ps[i]) << std::endl;
    return true;
}

/// makeCircularTrajectory (Projtable.hxx:138-165): n_proj views on a circle about the Y axis, normalised matrices.
inline std::vector<Geometry::ProjectionMatrix> makeCircularTrajectory(int n_proj, double sid, double sdd, int n_u, int n_v,
                                                                      double max_angle_deg, double pixel_spacing)
{
    std::vector<double> raw((size_t)n_proj * 12);
    ecc_make_circular_trajectory(n_proj, sid, sdd, n_u, n_v, max_angle_deg, pixel_spacing, raw.data());
    std::vector<Geometry::ProjectionMatrix> Ps(n_proj);
    for (int i = 0; i < n_proj; i++)
        for (int k = 0; k < 12; k++) Ps[i].data()[k] = raw[(size_t)i * 12 + k];
    return Ps;
}

}  // namespace ProjTable

#endif
