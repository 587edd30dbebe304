// picsp_b200/csrc/host/host.hpp — C++ host driver above the kernel ABI (no CUDA in here).
#pragma once
#include <cstdint>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../../../include/picsp_b200.h"
#include "../../../include/picsp_b200_host.h"

namespace picsp_host {

// ---- INI reader with iniparser 3.1's rules (reference lib/iniparser/src/iniparser.c:555-613, 408-437)
class IniFile {
public:
    bool load(const std::string &path, std::string *err);
    bool has(const std::string &key) const;
    int get_int(const std::string &key, int notfound) const;          // strtol(.., 0): "1e9" parses as 1
    double get_double(const std::string &key, double notfound) const; // atof
    const std::map<std::string, std::string> &entries() const { return kv_; }
private:
    std::map<std::string, std::string> kv_;   // "section:key" (lower case) -> value
};

int parse_run_config(const std::string &path, picsp_run_config &cfg, bool banner, std::string *err);

// ---- loader (src/main.cpp:567-640)
struct Loader {
    std::mt19937 gen;
    std::uniform_real_distribution<double> dist{0.0, 1.0};
    double x_carry = 0.0;    // loadType 2: the stale `x` the reference's self-referencing initialiser reads (main.cpp:599)
    explicit Loader(uint32_t seed) : gen(seed) {}
    double rnd() { return dist(gen); }
    void fill(const picsp_run_config &cfg, int species, double *x, double *y, double *vx, double *vy);
};

// ---- minimal HDF5 writer: exactly the subset picsp's output uses (src/main.cpp:21-36, 1142-1247)
class H5Writer {
public:
    // shadow = true: another process (rank 0 of a sharded run) owns the file's metadata; this writer only tracks the
    // same sequence of dataset offsets and writes row slices into the existing file
    bool open(const std::string &path, std::string *err, bool shadow = false);
    void create_group(const std::string &abs_name);                                      // "/particle.e"
    void write_dataset_f64(const std::string &abs_name, const double *data, uint64_t d0, uint64_t d1);
    // sharded runs: reserve the raw-data block of a dataset (every rank calls it in the same order and gets the same
    // address), then every rank writes its own rows of it
    uint64_t reserve_dataset_f64(const std::string &abs_name, uint64_t d0, uint64_t d1);
    bool write_rows(uint64_t data_addr, uint64_t row_lo, uint64_t nrows, uint64_t d1, const double *data);
    void write_attr_f64(const std::string &name, double v);                               // root attributes
    void write_attr_i32(const std::string &name, int32_t v);
    bool close(std::string *err);
private:
    struct Obj { std::string name; uint64_t header_addr = 0; bool is_group = false; std::vector<size_t> children; 
                 uint64_t data_addr = 0, d0 = 0, d1 = 0; };
    struct Attr { std::string name; bool is_int; double f; int32_t i; };
    std::vector<Obj> objs_;       // objs_[0] = root
    std::vector<Attr> attrs_;
    FILE *fp_ = nullptr;
    uint64_t eof_ = 0;
    bool shadow_ = false;
    size_t find_or_make_group(const std::string &abs_name);
    uint64_t append(const std::vector<uint8_t> &bytes);
    uint64_t write_group(size_t idx, uint16_t leaf_k);
};

int run(const std::string &ini_path, const std::string &out_path, int max_steps, bool quiet, int device, std::string *err,
        int rank = 0, int nranks = 1);

}  // namespace picsp_host
