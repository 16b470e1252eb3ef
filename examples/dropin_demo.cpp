// dropin_demo.cpp -- a PDWT client written against the reference's class interface (reference src/wt.h:20-76; the
// call sequence is the README's / demo.cpp's: forward -> norm1 -> soft_threshold -> norm1 -> inverse -> get_image),
// compiled against include/wt.h and linked with libpdwt_b200.so instead of libpdwt.so.  Nothing here knows about
// the C ABI underneath.
//
//   g++ -O2 -Iinclude -I/usr/local/cuda/include examples/dropin_demo.cpp -Lpdwt_b200 -lpdwt_b200 \
//       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/pdwt_b200 -o dropin_demo
//   ./dropin_demo <Nr> <Nc> <wavelet> <levels> <separable> <swt> <beta> <in.f32> <out.f32>
// Prints "norm1_before norm1_after max_abs_reconstruction_change" and writes the reconstructed image.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "wt.h"

int main(int argc, char** argv)
{
    if (argc < 10) {
        fprintf(stderr, "usage: %s Nr Nc wavelet levels separable swt beta in.f32 out.f32\n", argv[0]);
        return 2;
    }
    const int Nr = atoi(argv[1]), Nc = atoi(argv[2]), levels = atoi(argv[4]), sep = atoi(argv[5]), swt = atoi(argv[6]);
    const float beta = (float)atof(argv[7]);
    const size_t n = (size_t)Nr * Nc;
    float* img = (float*)malloc(n * sizeof(float));
    FILE* f = fopen(argv[8], "rb");
    if (!img || !f || fread(img, sizeof(float), n, f) != n) {
        fprintf(stderr, "cannot read %s\n", argv[8]);
        return 1;
    }
    fclose(f);

    Wavelets W(img, Nr, Nc, argv[3], levels, 1, sep, 0, swt);   // same argument order as the reference constructor
    if (W.state == W_CREATION_ERROR) {
        fprintf(stderr, "creation error\n");
        return 1;
    }
    W.print_informations();
    W.forward();
    const float n0 = W.norm1();
    W.soft_threshold(beta, 0, 0);
    const float n1 = W.norm1();
    W.inverse();
    float* out = (float*)malloc(n * sizeof(float));
    if ((size_t)W.get_image(out) != n) return 1;
    float worst = 0.f;
    for (size_t i = 0; i < n; i++) worst = fmaxf(worst, fabsf(out[i] - img[i]));
    printf("RESULT %.9g %.9g %.9g\n", n0, n1, worst);
    f = fopen(argv[9], "wb");
    fwrite(out, sizeof(float), n, f);
    fclose(f);
    free(img);
    free(out);
    return W.state == W_INVERSE ? 0 : 1;
}
