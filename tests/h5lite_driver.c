/* Exercises host/standins/h5lite.c through the HDF5 API subset the reference uses: reopen-per-save cadence, groups with
 * attributes, compound datasets written as two hyperslabs, padded memory spaces, H5LTmake_dataset.  Built and checked by
 * tests/test_hdf5_output.py with the independent reader tests/h5parse.py. */
#include "hdf5.h"
#include "hdf5_hl.h"
#include <stdio.h>
#include <stdlib.h>
typedef struct { double re, im; } cx;
int main(int argc, char** argv) {
    const char* path = argv[1];
    int ngroups = atoi(argv[2]);
    hid_t f = H5Fcreate(path, H5F_ACC_TRUNC, H5P_DEFAULT, H5P_DEFAULT);
    hid_t ct = H5Tcreate(H5T_COMPOUND, sizeof(cx));
    H5Tinsert(ct, "r", 0, H5T_NATIVE_DOUBLE); H5Tinsert(ct, "i", 8, H5T_NATIVE_DOUBLE);
    H5Fclose(f);
    for (int it = 0; it < ngroups; ++it) {
        f = H5Fopen(path, H5F_ACC_RDWR, H5P_DEFAULT);
        char name[64]; sprintf(name, "/Iter_%05d", it);
        if (H5Lexists(f, name, H5P_DEFAULT)) return 2;
        hid_t g = H5Gcreate(f, name, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        hsize_t one = 1; hid_t as = H5Screate_simple(1, &one, NULL);
        double t = 0.5 * it, dt = 1e-3;
        hid_t a = H5Acreate(g, "TimeValue", H5T_NATIVE_DOUBLE, as, H5P_DEFAULT, H5P_DEFAULT); H5Awrite(a, H5T_NATIVE_DOUBLE, &t); H5Aclose(a);
        a = H5Acreate(g, "TimeStep", H5T_NATIVE_DOUBLE, as, H5P_DEFAULT, H5P_DEFAULT); H5Awrite(a, H5T_NATIVE_DOUBLE, &dt); H5Aclose(a);
        H5Sclose(as);
        /* complex field [4][3][5][3], written as two x-slabs from a buffer */
        hsize_t dims[4] = {4, 3, 5, 3}, slab[4] = {2, 3, 5, 3};
        hid_t ds = H5Screate_simple(4, dims, NULL);
        hid_t d = H5Dcreate(g, "u_hat", ct, ds, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        for (int half = 0; half < 2; ++half) {
            cx buf[2 * 3 * 5 * 3];
            for (int e = 0; e < 90; ++e) { buf[e].re = it + 0.001 * (half * 90 + e); buf[e].im = -(double)(half * 90 + e); }
            hid_t ms = H5Screate_simple(4, slab, NULL);
            hsize_t z[4] = {0, 0, 0, 0}, off[4] = {2 * (hsize_t)half, 0, 0, 0};
            H5Sselect_hyperslab(ms, H5S_SELECT_SET, z, NULL, slab, NULL);
            H5Sselect_hyperslab(ds, H5S_SELECT_SET, off, NULL, slab, NULL);
            if (H5Dwrite(d, ct, ms, ds, H5P_DEFAULT, buf) < 0) return 3;
            H5Sclose(ms);
        }
        H5Dclose(d); H5Sclose(ds);
        /* real field with padded memory rows: file [2][4][3], memory [2][6][3] */
        hsize_t rd[3] = {2, 4, 3}, md[3] = {2, 6, 3}, z3[3] = {0, 0, 0};
        double rbuf[2 * 6 * 3];
        for (int e = 0; e < 36; ++e) rbuf[e] = e + 100 * it;
        hid_t rs = H5Screate_simple(3, rd, NULL), ms = H5Screate_simple(3, md, NULL);
        H5Sselect_hyperslab(ms, H5S_SELECT_SET, z3, NULL, rd, NULL);
        H5Sselect_hyperslab(rs, H5S_SELECT_SET, z3, NULL, rd, NULL);
        d = H5Dcreate(g, "u", H5T_NATIVE_DOUBLE, rs, H5P_DEFAULT, H5P_DEFAULT, H5P_DEFAULT);
        if (H5Dwrite(d, H5T_NATIVE_DOUBLE, ms, rs, H5P_DEFAULT, rbuf) < 0) return 4;
        H5Dclose(d); H5Sclose(rs); H5Sclose(ms);
        H5Gclose(g);
        if (H5Fclose(f) < 0) return 5;
    }
    f = H5Fopen(path, H5F_ACC_RDWR, H5P_DEFAULT);
    int k[7] = {0, 1, 2, 3, -3, -2, -1};
    hsize_t n7 = 7;
    H5LTmake_dataset(f, "kx", 1, &n7, H5T_NATIVE_INT, k);
    double tt[3] = {0.0, 0.5, 1.0}; hsize_t n3 = 3;
    H5LTmake_dataset(f, "Time", 1, &n3, H5T_NATIVE_DOUBLE, tt);
    H5Fclose(f);
    H5Tclose(ct);
    return 0;
}
