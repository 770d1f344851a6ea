/* hdf5_hl.h -- H5LTmake_dataset, the one "lite" call the reference uses (hdf5_funcs.c:1058-1165); see hdf5.h. */
#ifndef NSB200_H5LITE_HDF5_HL_H
#define NSB200_H5LITE_HDF5_HL_H
#include "hdf5.h"
#ifdef __cplusplus
extern "C" {
#endif
herr_t H5LTmake_dataset(hid_t loc, const char* name, int rank, const hsize_t* dims, hid_t type, const void* buffer);
#ifdef __cplusplus
}
#endif
#endif
