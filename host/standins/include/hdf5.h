/* hdf5.h -- the subset of the HDF5 C API that the reference's hdf5_funcs.c calls (35 names, listed in SURVEY.md 8c),
 * implemented by host/standins/h5lite.c as a native writer of real HDF5 files (superblock v0, symbol-table groups,
 * version-1 object headers, contiguous datasets, compound / IEEE f64 / i32 datatypes, attributes).  libhdf5 is absent
 * from this image; with this header the reference's OWN hdf5_funcs.c compiles unmodified and produces its own file
 * layout (hdf5_funcs.c:172-206 datasets, :1254-1279 attributes, :1058-1165 end-of-run series).  A deployment that has
 * libhdf5 simply puts the real <hdf5.h> first on the include path.  Single process: the MPI-IO property calls are
 * accepted and ignored.  Plus the read-side calls the restart hook uses (H5Dopen, H5Dread, H5Dget_space,
 * H5Sget_simple_extent_ndims / _dims, H5Aopen, H5Aread) on files opened with H5F_ACC_RDONLY. */
#ifndef NSB200_H5LITE_HDF5_H
#define NSB200_H5LITE_HDF5_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hid_t;
typedef int herr_t;
typedef int htri_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;

#define H5P_DEFAULT ((hid_t)0)
#define H5S_ALL ((hid_t)0)
#define H5F_ACC_RDONLY 0x0000u
#define H5F_ACC_RDWR 0x0001u
#define H5F_ACC_TRUNC 0x0002u
typedef enum { H5S_SELECT_SET = 0 } H5S_seloper_t;
typedef enum { H5T_COMPOUND = 6 } H5T_class_t;
typedef enum { H5FD_MPIO_INDEPENDENT = 0, H5FD_MPIO_COLLECTIVE = 1 } H5FD_mpio_xfer_t;
/* property list classes */
#define H5P_FILE_ACCESS ((hid_t)1)
#define H5P_DATASET_XFER ((hid_t)2)
/* predefined native datatypes (x86-64 / aarch64 little endian) */
#define H5T_NATIVE_DOUBLE ((hid_t)11)
#define H5T_NATIVE_INT ((hid_t)12)

hid_t H5Fcreate(const char* name, unsigned flags, hid_t fcpl, hid_t fapl);
hid_t H5Fopen(const char* name, unsigned flags, hid_t fapl);
herr_t H5Fclose(hid_t file);

hid_t H5Gcreate(hid_t loc, const char* name, hid_t lcpl, hid_t gcpl, hid_t gapl);
hid_t H5Gopen(hid_t loc, const char* name, hid_t gapl);
herr_t H5Gclose(hid_t group);
htri_t H5Lexists(hid_t loc, const char* name, hid_t lapl);

hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t* maxdims);
herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t op, const hsize_t* start, const hsize_t* stride, const hsize_t* count,
                           const hsize_t* block);
herr_t H5Sclose(hid_t space);

hid_t H5Tcreate(H5T_class_t cls, size_t size);
herr_t H5Tinsert(hid_t type, const char* name, size_t offset, hid_t member);
herr_t H5Tclose(hid_t type);

hid_t H5Dcreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl);
herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t dxpl, const void* buf);
/* read side (restart from a saved state, host/nsb200_hooks.c; not called by the reference's hdf5_funcs.c) */
hid_t H5Dopen(hid_t loc, const char* name, hid_t dapl);
herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t dxpl, void* buf);
hid_t H5Dget_space(hid_t dset);
int H5Sget_simple_extent_ndims(hid_t space);
int H5Sget_simple_extent_dims(hid_t space, hsize_t* dims, hsize_t* maxdims);
herr_t H5Dclose(hid_t dset);

hid_t H5Acreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t acpl, hid_t aapl);
herr_t H5Awrite(hid_t attr, hid_t mem_type, const void* buf);
hid_t H5Aopen(hid_t loc, const char* name, hid_t aapl);      /* read side */
herr_t H5Aread(hid_t attr, hid_t mem_type, void* buf);
herr_t H5Aclose(hid_t attr);

hid_t H5Pcreate(hid_t cls);
herr_t H5Pclose(hid_t plist);
/* MPI-IO: accepted and ignored (single process).  Declared with loose types so that both the stand-in mpi.h and a
 * real one work. */
#define H5Pset_fapl_mpio(plist, comm, info) h5lite_noop((plist))
#define H5Pset_dxpl_mpio(plist, mode) h5lite_noop((plist))
herr_t h5lite_noop(hid_t plist);

#ifdef __cplusplus
}
#endif
#endif
