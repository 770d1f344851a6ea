/* Read side of host/standins/h5lite.c: opens, in ANOTHER process, the file tests/h5lite_driver.c wrote and checks what the
 * restart hook relies on - read-only open of a foreign file, H5Lexists walk to the last /Iter_%05d group, H5Dopen /
 * H5Dget_space / H5Dread of the compound u_hat (whole and as an x-slab hyperslab into a contiguous buffer), plain f64 and
 * i32 datasets - against the values the driver generated.  Built and run by tests/test_hdf5_output.py. */
#include "hdf5.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
typedef struct { double re, im; } cx;
int main(int argc, char** argv) {
    if (argc < 3) return 100;
    const char* path = argv[1];
    const int ngroups = atoi(argv[2]);
    if (H5Fopen(path, H5F_ACC_RDWR, H5P_DEFAULT) >= 0) return 1;             /* foreign files are read-only */
    if (H5Fopen("/nonexistent/file.h5", H5F_ACC_RDONLY, H5P_DEFAULT) >= 0) return 2;
    hid_t f = H5Fopen(path, H5F_ACC_RDONLY, H5P_DEFAULT);
    if (f < 0) return 3;
    int last = -1;
    char name[64];
    for (int i = 0;; ++i) {
        snprintf(name, sizeof name, "Iter_%05d", i);
        if (H5Lexists(f, name, H5P_DEFAULT) > 0) last = i; else break;
    }
    if (last != ngroups - 1) return 4;
    if (H5Gcreate(f, "/New", H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT) >= 0) return 5;   /* no writes */
    hid_t ct = H5Tcreate(H5T_COMPOUND, sizeof(cx));
    H5Tinsert(ct, "r", 0, H5T_NATIVE_DOUBLE); H5Tinsert(ct, "i", 8, H5T_NATIVE_DOUBLE);
    snprintf(name, sizeof name, "/Iter_%05d/u_hat", last);
    hid_t d = H5Dopen(f, name, H5P_DEFAULT);
    if (d < 0) return 6;
    hid_t fs = H5Dget_space(d);
    hsize_t dims[4] = {0, 0, 0, 0};
    if (H5Sget_simple_extent_ndims(fs) != 4 || H5Sget_simple_extent_dims(fs, dims, NULL) != 4) return 7;
    if (dims[0] != 4 || dims[1] != 3 || dims[2] != 5 || dims[3] != 3) return 8;
    cx all[180];
    if (H5Dread(d, ct, H5S_ALL, H5S_ALL, H5P_DEFAULT, all) < 0) return 9;
    for (int e = 0; e < 180; ++e)
        if (fabs(all[e].re - (last + 0.001 * e)) > 1e-15 * (1 + last) || all[e].im != -(double)e) return 10;
    /* the second x-slab only, as a rank reading its own planes does */
    hsize_t slab[4] = {2, 3, 5, 3}, off[4] = {2, 0, 0, 0};
    hid_t ms = H5Screate_simple(4, slab, NULL);
    if (H5Sselect_hyperslab(fs, H5S_SELECT_SET, off, NULL, slab, NULL) < 0) return 11;
    cx half[90];
    if (H5Dread(d, ct, ms, fs, H5P_DEFAULT, half) < 0) return 12;
    for (int e = 0; e < 90; ++e)
        if (half[e].re != all[90 + e].re || half[e].im != all[90 + e].im) return 13;
    double wrong[180];
    if (H5Dread(d, H5T_NATIVE_DOUBLE, H5S_ALL, H5S_ALL, H5P_DEFAULT, wrong) >= 0) return 14;   /* no type conversion */
    H5Sclose(ms); H5Sclose(fs); H5Dclose(d);
    snprintf(name, sizeof name, "/Iter_%05d/u", last);
    d = H5Dopen(f, name, H5P_DEFAULT);
    double u[24];
    if (d < 0 || H5Dread(d, H5T_NATIVE_DOUBLE, H5S_ALL, H5S_ALL, H5P_DEFAULT, u) < 0) return 15;
    for (int a = 0; a < 2; ++a) for (int b = 0; b < 4; ++b) for (int c = 0; c < 3; ++c)
        if (u[(a * 4 + b) * 3 + c] != (a * 6 + b) * 3 + c + 100 * last) return 16;          /* file [2][4][3] from memory [2][6][3] */
    H5Dclose(d);
    d = H5Dopen(f, "kx", H5P_DEFAULT);
    int k[7];
    if (d < 0 || H5Dread(d, H5T_NATIVE_INT, H5S_ALL, H5S_ALL, H5P_DEFAULT, k) < 0) return 17;
    if (k[0] != 0 || k[3] != 3 || k[4] != -3 || k[6] != -1) return 18;
    H5Dclose(d);
    if (H5Dopen(f, "/Iter_00000", H5P_DEFAULT) >= 0) return 19;                            /* a group is not a dataset */
    snprintf(name, sizeof name, "/Iter_%05d", last);
    hid_t g = H5Gopen(f, name, H5P_DEFAULT);
    hid_t a = H5Aopen(g, "TimeValue", H5P_DEFAULT);
    double tv = -1.0, ts = -1.0;
    if (g < 0 || a < 0 || H5Aread(a, H5T_NATIVE_DOUBLE, &tv) < 0 || tv != 0.5 * last) return 21;
    H5Aclose(a);
    a = H5Aopen(g, "TimeStep", H5P_DEFAULT);
    if (a < 0 || H5Aread(a, H5T_NATIVE_DOUBLE, &ts) < 0 || ts != 1e-3) return 22;
    H5Aclose(a);
    if (H5Aopen(g, "NoSuchAttribute", H5P_DEFAULT) >= 0) return 23;
    H5Gclose(g);
    H5Tclose(ct);
    if (H5Fclose(f) < 0) return 20;
    return 0;
}
