/* oracle/ref_gpu_harness.cpp — TEST INFRASTRUCTURE ONLY (builds oracle/_ref/libpicsp_ref_gpu*.so).
 *
 * The boundary proof: the reference's own translation unit with the maintainer's patch of INTEGRATION.md
 * section 2 applied (oracle/_ref/main_gpu.cpp, generated from /root/reference/src/main.cpp by oracle/gpu_patch.py;
 * never committed), compiled against the same shim headers as the CPU oracle and LINKED AGAINST
 * picsp_b200/libpicsp_b200.so.  Its main() is the reference's main(): parse_ini_file, init, banner, writeSpecies,
 * writePot, writeKE are the reference's own code; only the hot-path calls go through the C ABI.  The H5 shim
 * captures what it writes, exactly as for the unmodified TU (ref_harness.cpp).
 */
#define main picsp_refgpu_main_entry
#include PICSP_PATCHED_MAIN_CPP   /* -DPICSP_PATCHED_MAIN_CPP="\".../oracle/_ref/main_gpu.cpp\"" */
#undef main

extern "C" {

int picsp_refgpu_main(const char *ini_path) {
    /* re-run the reference's static initialisers (main.cpp:29-36): its main() deletes them on exit */
    file = new H5File(FILE_NAME, H5F_ACC_TRUNC);
    groupE = new Group(file->createGroup("/particle.e"));
    groupI = new Group(file->createGroup("/particle.i"));
    groupT = new Group(file->createGroup("/timedata"));
    groupP = new Group(file->createGroup("/phi"));
    groupDE = new Group(file->createGroup("/den.e"));
    groupDI = new Group(file->createGroup("/den.i"));
    mt_gen.seed(0); rnd_dist.reset();
    std::string prog("picsp"), p(ini_path);
    char *argv[3] = {&prog[0], &p[0], nullptr};
    return picsp_refgpu_main_entry(2, argv);
}

long picsp_ref_h5_count(void) { return (long)picsp_shim_h5_registry().size(); }
const char *picsp_ref_h5_name(long i) { return picsp_shim_h5_registry()[i].name.c_str(); }
void picsp_ref_h5_meta(long i, long long *meta6) {
    const picsp_shim_h5_record &r = picsp_shim_h5_registry()[i];
    meta6[0] = r.is_attr; meta6[1] = r.elem; meta6[2] = r.rank;
    meta6[3] = (long long)r.dims[0]; meta6[4] = (long long)r.dims[1]; meta6[5] = (long long)r.bytes.size();
}
const void *picsp_ref_h5_data(long i) { return picsp_shim_h5_registry()[i].bytes.data(); }
long picsp_ref_h5_group_count(void) { return (long)picsp_shim_h5_groups().size(); }
const char *picsp_ref_h5_group_name(long i) { return picsp_shim_h5_groups()[i].c_str(); }

}  /* extern "C" */
