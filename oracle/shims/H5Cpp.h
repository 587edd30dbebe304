/* oracle/shims/H5Cpp.h — TEST INFRASTRUCTURE ONLY.
 *
 * In-memory stand-in for the handful of HDF5 C++ API types the reference uses
 * (/root/reference/src/main.cpp:24-36, 195-199, 1142-1247) so the UNMODIFIED
 * reference translation unit compiles without libhdf5.  HDF5 does no arithmetic
 * on this path; the shim only records what the reference writes (object name,
 * element type, shape, raw bytes) in a process-wide registry that the harness
 * exposes through `picsp_ref_h5_*`, which is how whole-run golden vectors
 * (`/timedata/energy`, `/phi/<ts>`, ...) are captured from the reference's own
 * `main()`.
 */
#ifndef PICSP_ORACLE_H5CPP_SHIM_H
#define PICSP_ORACLE_H5CPP_SHIM_H

#include <string>
#include <vector>
#include <cstring>
#include <cstddef>

typedef std::string H5std_string;
typedef unsigned long long hsize_t;

#define H5F_ACC_TRUNC 2u
enum picsp_shim_h5s_class { H5S_SCALAR = 0 };

struct picsp_shim_h5_record {
    std::string name;
    int is_attr;          /* 1 = root attribute, 0 = dataset */
    int elem;             /* 0 = f64, 1 = i32 */
    int rank;
    hsize_t dims[2];
    std::vector<unsigned char> bytes;
};

inline std::vector<picsp_shim_h5_record> &picsp_shim_h5_registry() {
    static std::vector<picsp_shim_h5_record> *reg = new std::vector<picsp_shim_h5_record>();
    return *reg;
}
inline std::vector<std::string> &picsp_shim_h5_groups() {
    static std::vector<std::string> *g = new std::vector<std::string>();
    return *g;
}

namespace H5 {

struct PredType {
    int code;
    static const PredType NATIVE_DOUBLE;
    static const PredType NATIVE_INT;
};
inline const PredType PredType::NATIVE_DOUBLE = {0};
inline const PredType PredType::NATIVE_INT = {1};

class DataType {
public:
    int code;
    DataType(const PredType &p) : code(p.code) {}
    size_t size() const { return code == 0 ? 8 : 4; }
};

class DataSpace {
public:
    int rank;
    hsize_t dims[2];
    DataSpace(int r, const hsize_t *d) : rank(r) { dims[0] = r > 0 ? d[0] : 1; dims[1] = r > 1 ? d[1] : 1; }
    DataSpace(picsp_shim_h5s_class) : rank(0) { dims[0] = dims[1] = 1; }
    size_t count() const { return (size_t)dims[0] * (size_t)dims[1]; }
};

class Group {
public:
    std::string name;
};

class DataSet {
public:
    size_t rec;
    void write(const void *buf, const DataType &t) {
        picsp_shim_h5_record &r = picsp_shim_h5_registry()[rec];
        size_t n = (size_t)r.dims[0] * (size_t)r.dims[1] * t.size();
        r.bytes.resize(n);
        std::memcpy(r.bytes.data(), buf, n);
    }
};

class Attribute {
public:
    size_t rec;
    void write(const DataType &t, const void *buf) {
        picsp_shim_h5_record &r = picsp_shim_h5_registry()[rec];
        r.bytes.resize(t.size());
        std::memcpy(r.bytes.data(), buf, t.size());
    }
};

class H5File {
public:
    std::string path;
    H5File(const H5std_string &name, unsigned) : path(name) {
        /* H5F_ACC_TRUNC: a new file starts empty */
        picsp_shim_h5_registry().clear();
        picsp_shim_h5_groups().clear();
    }
    Group createGroup(const H5std_string &name) {
        picsp_shim_h5_groups().push_back(name);
        Group g; g.name = name; return g;
    }
    DataSet createDataSet(const H5std_string &name, const DataType &t, const DataSpace &s) {
        picsp_shim_h5_record r;
        r.name = name; r.is_attr = 0; r.elem = t.code; r.rank = s.rank;
        r.dims[0] = s.dims[0]; r.dims[1] = s.dims[1];
        picsp_shim_h5_registry().push_back(r);
        DataSet d; d.rec = picsp_shim_h5_registry().size() - 1; return d;
    }
    Attribute createAttribute(const H5std_string &name, const DataType &t, const DataSpace &s) {
        picsp_shim_h5_record r;
        r.name = name; r.is_attr = 1; r.elem = t.code; r.rank = s.rank;
        r.dims[0] = s.dims[0]; r.dims[1] = s.dims[1];
        picsp_shim_h5_registry().push_back(r);
        Attribute a; a.rec = picsp_shim_h5_registry().size() - 1; return a;
    }
};

} /* namespace H5 */

#endif
