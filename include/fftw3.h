/* pnfft-b200: placeholder for <fftw3.h>.  The reference's test drivers include it (tests/simple_test.c:6) but use no
 * FFTW symbol; the oversampled FFT of this library runs on the device.  Nothing is declared here on purpose. */
#ifndef PNFFT_B200_FFTW3_H
#define PNFFT_B200_FFTW3_H 1
#endif
