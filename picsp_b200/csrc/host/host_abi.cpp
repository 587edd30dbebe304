// extern "C" face of the host driver (include/picsp_b200_host.h).
#include <algorithm>
#include <cstring>
#include <iostream>

#include "host.hpp"

struct picsp_loader { picsp_host::Loader impl; explicit picsp_loader(uint32_t s) : impl(s) {} };

extern "C" {

int picsp_host_parse_ini(const char *path, picsp_run_config *out, int print_banner) {
    if (!path || !out) return PICSP_ERR_INVALID;
    std::string err;
    try { return picsp_host::parse_run_config(path, *out, print_banner != 0, &err); }
    catch (...) { return PICSP_ERR_INVALID; }
}

int64_t picsp_host_ini_dump(const char *path, char *buf, int64_t buflen) {
    if (!path) return PICSP_ERR_INVALID;
    picsp_host::IniFile ini;
    std::string err;
    try { if (!ini.load(path, &err)) return PICSP_ERR_INVALID; } catch (...) { return PICSP_ERR_INVALID; }
    std::string out;
    for (const auto &kv : ini.entries()) { out += kv.first; out += '\t'; out += kv.second; out += '\n'; }
    if (buf && buflen > 0) {
        const size_t n = std::min<size_t>(out.size(), (size_t)buflen - 1);
        std::memcpy(buf, out.data(), n); buf[n] = 0;
    }
    return (int64_t)out.size() + 1;
}

picsp_loader *picsp_host_loader_create(uint32_t seed) {
    try { return new picsp_loader(seed); } catch (...) { return nullptr; }
}
void picsp_host_loader_destroy(picsp_loader *ld) { delete ld; }

int picsp_host_loader_fill(picsp_loader *ld, const picsp_run_config *cfg, int species,
                           double *x, double *y, double *vx, double *vy) {
    if (!ld || !cfg || (species != 0 && species != 1) || !x || !y || !vx || !vy) return PICSP_ERR_INVALID;
    try { ld->impl.fill(*cfg, species, x, y, vx, vy); } catch (...) { return PICSP_ERR_INVALID; }
    return PICSP_OK;
}

int picsp_host_run(const char *ini_path, const char *out_path, int max_steps, int quiet, int device) {
    if (!ini_path) return PICSP_ERR_INVALID;
    std::string err;
    int rc;
    try { rc = picsp_host::run(ini_path, out_path ? out_path : "", max_steps, quiet != 0, device, &err); }
    catch (const std::exception &e) { err = e.what(); rc = PICSP_ERR_INVALID; }
    if (rc != PICSP_OK && !err.empty()) std::cerr << "picsp_b200: " << err << std::endl;
    return rc;
}

int picsp_host_run_ranked(const char *ini_path, const char *out_path, int max_steps, int quiet, int device, int rank, int nranks) {
    if (!ini_path || nranks < 1 || rank < 0 || rank >= nranks) return PICSP_ERR_INVALID;
    std::string err;
    int rc;
    try { rc = picsp_host::run(ini_path, out_path ? out_path : "", max_steps, quiet != 0, device, &err, rank, nranks); }
    catch (const std::exception &e) { err = e.what(); rc = PICSP_ERR_INVALID; }
    if (rc != PICSP_OK && !err.empty()) std::cerr << "picsp_b200 (rank " << rank << "): " << err << std::endl;
    return rc;
}

picsp_h5 *picsp_host_h5_open(const char *path) {
    if (!path) return nullptr;
    auto *w = new picsp_host::H5Writer();
    std::string err;
    if (!w->open(path, &err)) { delete w; return nullptr; }
    return reinterpret_cast<picsp_h5 *>(w);
}
int picsp_host_h5_group(picsp_h5 *h, const char *n) {
    if (!h || !n) return PICSP_ERR_INVALID;
    reinterpret_cast<picsp_host::H5Writer *>(h)->create_group(n); return PICSP_OK;
}
int picsp_host_h5_dataset_f64(picsp_h5 *h, const char *n, const double *data, uint64_t d0, uint64_t d1) {
    if (!h || !n || (!data && d0 != 0 && d1 != 0)) return PICSP_ERR_INVALID;
    reinterpret_cast<picsp_host::H5Writer *>(h)->write_dataset_f64(n, data, d0, d1); return PICSP_OK;
}
int picsp_host_h5_attr_f64(picsp_h5 *h, const char *n, double v) {
    if (!h || !n) return PICSP_ERR_INVALID;
    reinterpret_cast<picsp_host::H5Writer *>(h)->write_attr_f64(n, v); return PICSP_OK;
}
int picsp_host_h5_attr_i32(picsp_h5 *h, const char *n, int32_t v) {
    if (!h || !n) return PICSP_ERR_INVALID;
    reinterpret_cast<picsp_host::H5Writer *>(h)->write_attr_i32(n, v); return PICSP_OK;
}
int picsp_host_h5_close(picsp_h5 *h) {
    if (!h) return PICSP_ERR_INVALID;
    auto *w = reinterpret_cast<picsp_host::H5Writer *>(h);
    std::string err;
    bool ok = w->close(&err);
    delete w;
    return ok ? PICSP_OK : PICSP_ERR_INVALID;
}

}  // extern "C"
