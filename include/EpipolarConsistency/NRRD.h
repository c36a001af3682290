// NRRD.h -- minimal NRRD::ImageView<T> / NRRD::Image<T> for the facade: the container type the reference's
// RadonIntermediate API exposes (HeaderOnly/NRRD/nrrd_image_view.hxx:14-263, nrrd_image.hxx:12) and the
// on-disk layout of a saved Radon intermediate (HeaderOnly/NRRD/nrrd.hxx:132-183 save, :188-245 header parse):
//   "NRRD0004", fields "name: value" in lexical order, keys "name:=value" in lexical order, the comment
//   "# Offset to raw data: %8d bytes.", an empty line, raw little-endian data, x fastest.
// Files written here load in the reference's tools and vice versa.  Only what the hot path
// needs is implemented: 2-D/3-D float/double/int images, raw encoding.
#ifndef ECC_FACADE_NRRD_H
#define ECC_FACADE_NRRD_H

#include <cstdio>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

template <typename T> inline std::string toString(const T& in)
{
    std::ostringstream s;
    s << in;  // default stream formatting: 6 significant digits, as HeaderOnly/NRRD/StringConversion.hxx:33-38
    return s.str();
}
template <typename T> inline T stringTo(const std::string& in)
{
    T v = T();
    std::istringstream s(in);
    s >> v;
    return v;
}
template <> inline std::string stringTo<std::string>(const std::string& in) { return in; }

namespace NRRD {

template <typename T> struct TypeName;
template <> struct TypeName<float> { static const char* name() { return "float"; } };
template <> struct TypeName<double> { static const char* name() { return "double"; } };
template <> struct TypeName<int> { static const char* name() { return "int"; } };
template <> struct TypeName<unsigned char> { static const char* name() { return "unsigned char"; } };

template <typename T> class ImageView {
protected:
    std::vector<int> dim;
    std::vector<double> element_spacing;
    T* data;

    static void clampCell(int& i, double& f, int n)
    {
        if (i < 0) { i = 0; f = 0; }
        if (i > n - 2) { f = 1.0; i = n - 2; }
    }

public:
    std::map<std::string, std::string> meta_info;             // "key:=value" lines
    mutable std::map<std::string, std::string> nrrd_header;   // extra "field: value" lines

    ImageView() : data(0x0) {}
    ImageView(int w, int h, int d = 1, T* dt = 0x0) : data(0x0) { set(w, h, d, dt); }

    void set(int w, int h, int d = 1, T* dt = 0x0)
    {
        dim.clear();
        dim.push_back(w);
        dim.push_back(h);
        if (d > 1) dim.push_back(d);
        element_spacing.assign(dim.size(), 1.0);
        data = dt;
    }
    int dimension() const { return (int)dim.size(); }
    int size(int i) const { return i < (int)dim.size() ? dim[i] : 1; }
    int length() const
    {
        if (dim.empty()) return 0;
        int n = 1;
        for (int d : dim) n *= d;
        return n;
    }
    double spacing(int i) const { return i < (int)element_spacing.size() ? element_spacing[i] : 1.0; }
    double& spacing(int i) { return element_spacing[i]; }
    operator T*() { return data; }
    operator const T*() const { return data; }
    bool operator!() const { return data == 0x0 || length() <= 0; }
    T& pixel(int x, int y, int z = 0) { return data[x + (size_t)y * dim[0] + (size_t)z * dim[0] * dim[1]]; }
    const T& pixel(int x, int y, int z = 0) const { return data[x + (size_t)y * dim[0] + (size_t)z * dim[0] * dim[1]]; }

    /// Bilinear lookup at a continuous pixel position, edges clamped: the CPU sampling of the reference's views
    /// (HeaderOnly/NRRD/nrrd_image_view.hxx:159-166,189-210; 2-D images: the xy-bilinear branch).  A position left of pixel 0
    /// reads pixel 0, one right of pixel n-2 reads pixel n-1 (cell n-2 with weight 1).
    double operator()(double x, double y) const
    {
        int ix = (int)x, iy = (int)y;
        double fx = x - ix, fy = y - iy;
        clampCell(ix, fx, dim[0]);
        clampCell(iy, fy, dim[1]);
        if (fx == 0 && fy == 0) return pixel(ix, iy);
        return (1.0 - fy) * ((1.0 - fx) * pixel(ix, iy) + fx * pixel(ix + 1, iy)) + fy * ((1.0 - fx) * pixel(ix, iy + 1) + fx * pixel(ix + 1, iy + 1));
    }

    bool save(const std::string& path) const
    {
        if (!*this) return false;
        std::ofstream f(path.c_str(), std::ios::binary);
        if (!f || !f.good()) return false;
        std::map<std::string, std::string> fields = nrrd_header;
        fields["type"] = TypeName<T>::name();
        fields["dimension"] = toString((int)dim.size());
        std::string sizes, spacings;
        for (size_t i = 0; i < dim.size(); i++) {
            sizes += (i ? " " : "") + toString(dim[i]);
            spacings += (i ? " " : "") + toString(element_spacing[i]);
        }
        fields["sizes"] = sizes;
        fields["spacings"] = spacings;
        fields["encoding"] = "raw";
        fields["endian"] = "little";
        f << "NRRD0004\n";
        for (auto& kv : fields) f << kv.first << ": " << kv.second << std::endl;
        for (auto& kv : meta_info)
            if (kv.second.find('\n') == std::string::npos) f << kv.first << ":=" << kv.second << std::endl;
        f << "# Offset to raw data: ";
        const std::string tail = " bytes.\n\n";
        const int offset = (int)f.tellp() + (int)tail.length() + 8;
        std::ostringstream num;
        num << std::setw(8) << std::setfill(' ') << offset;
        f << num.str() << tail;
        f.write((const char*)data, sizeof(T) * (size_t)length());
        return f.good();
    }
};

template <typename T> class Image : public ImageView<T> {
    std::vector<T> storage;
    using ImageView<T>::dim;
    using ImageView<T>::data;

public:
    Image() {}
    Image(int w, int h, int d = 1) { set(w, h, d); }
    explicit Image(const std::string& path) { load(path); }
    Image(const Image& o) : ImageView<T>() { clone(o); }
    Image& operator=(const Image& o)
    {
        if (this != &o) clone(o);
        return *this;
    }

    void set(int w, int h, int d = 1)
    {
        if (w <= 0 || h <= 0) {
            storage.clear();
            dim.clear();
            data = 0x0;
            return;
        }
        storage.assign((size_t)w * h * (d > 1 ? d : 1), T());
        ImageView<T>::set(w, h, d, storage.data());
    }
    // the reference's "set(0x0,0)": drop the data
    void set(const int*, int) { set(0, 0); }

    void clone(const ImageView<T>& o)
    {
        set(o.size(0), o.size(1), o.size(2));
        const T* src = o;
        if (src && data) std::copy(src, src + this->length(), data);
        this->meta_info = o.meta_info;
        this->nrrd_header = o.nrrd_header;
        for (int i = 0; i < this->dimension(); i++) this->element_spacing[i] = o.spacing(i);
    }

    bool load(const std::string& path)
    {
        set(0, 0);
        std::ifstream f(path.c_str(), std::ios::binary);
        if (!f || !f.good()) return false;
        std::string line;
        std::getline(f, line);
        if (line.compare(0, 4, "NRRD") != 0) return false;
        std::map<std::string, std::string> fields;
        this->meta_info.clear();
        while (std::getline(f, line)) {
            if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);
            if (line.empty()) break;  // raw data follows
            if (line[0] == '#') continue;
            const size_t c = line.find(':');
            if (c == std::string::npos || c + 1 >= line.size()) return false;
            const std::string key = line.substr(0, c), value = line.substr(c + 2 <= line.size() ? c + 2 : c + 1);
            if (line[c + 1] == '=') this->meta_info[key] = value;
            else if (line[c + 1] == ' ') fields[key] = value;
            else return false;
        }
        if (fields["encoding"] != "raw") return false;
        if (fields.count("endian") && fields["endian"] != "little") return false;
        std::vector<int> sz;
        {
            std::istringstream s(fields["sizes"]);
            int v;
            while (s >> v) sz.push_back(v);
        }
        if (sz.size() < 2 || sz.size() > 3) return false;
        set(sz[0], sz[1], sz.size() > 2 ? sz[2] : 1);
        if (fields.count("spacings")) {
            std::istringstream s(fields["spacings"]);
            double v;
            for (int i = 0; i < this->dimension() && (s >> v); i++) this->element_spacing[i] = v;
        }
        const std::string type = fields["type"];
        const size_t n = (size_t)this->length();
        if (type == TypeName<T>::name()) {
            f.read((char*)data, sizeof(T) * n);
        } else if (type == "float" || type == "double" || type == "int" || type == "unsigned char" || type == "uint8" || type == "uchar") {
            // convert on load like the reference's NRRD::load<T>
            if (type == "float") { std::vector<float> t(n); f.read((char*)t.data(), 4 * n); for (size_t i = 0; i < n; i++) data[i] = (T)t[i]; }
            else if (type == "double") { std::vector<double> t(n); f.read((char*)t.data(), 8 * n); for (size_t i = 0; i < n; i++) data[i] = (T)t[i]; }
            else if (type == "int") { std::vector<int> t(n); f.read((char*)t.data(), 4 * n); for (size_t i = 0; i < n; i++) data[i] = (T)t[i]; }
            else { std::vector<unsigned char> t(n); f.read((char*)t.data(), n); for (size_t i = 0; i < n; i++) data[i] = (T)t[i]; }
        } else {
            set(0, 0);
            return false;
        }
        if (!f) {
            set(0, 0);
            return false;
        }
        for (auto& kv : fields)
            if (kv.first != "type" && kv.first != "dimension" && kv.first != "sizes" && kv.first != "spacings" &&
                kv.first != "encoding" && kv.first != "endian")
                this->nrrd_header[kv.first] = kv.second;
        return true;
    }
};

}  // namespace NRRD

#endif
