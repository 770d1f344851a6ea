/* empty: gsl is included by data_types.h but unused by the Solver */
