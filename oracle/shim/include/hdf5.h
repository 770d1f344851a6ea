/* Type-only HDF5 stand-in so data_types.h compiles for the parity oracle; the reference's
 * hdf5_funcs.c is NOT compiled into oracle/_ref (oracle/shim/io_stub.c provides its three
 * entry points).  Test infrastructure, not product code. */
#ifndef NSB200_ORACLE_HDF5_H
#define NSB200_ORACLE_HDF5_H
#include <stdint.h>
typedef int64_t hid_t;
typedef int herr_t;
typedef unsigned long long hsize_t;
#endif
