/* see hdf5.h */
