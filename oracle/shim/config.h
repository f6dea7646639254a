/* TEST INFRASTRUCTURE ONLY -- the reference's autoconf-generated config.h is not
 * available; precision is selected with -DPNFFT_PREC_SINGLE on the command line
 * (oracle/Makefile).  Nothing else from config.h is used on the oracle path. */
#ifndef ORACLE_SHIM_CONFIG_H
#define ORACLE_SHIM_CONFIG_H 1
#endif
