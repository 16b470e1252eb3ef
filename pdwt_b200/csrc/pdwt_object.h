// pdwt_object.h -- the Wavelets object behind a Layer-B handle (internal: shared by pdwt_wavelets.cu and pdwt_sharded.cu)
#pragma once
#include "../../include/wt.h"

struct pdwt_wavelets {
    Wavelets W;
    pdwt_wavelets(const float* img, int Nr, int Nc, const char* wname, int levels, int memisonhost, int sep, int cs,
                  int swt, int ndim, int batch)
        : W(const_cast<float*>(img), Nr, Nc, wname, levels, memisonhost, sep, cs, swt, ndim, batch)
    {
    }
    explicit pdwt_wavelets(const Wavelets& o) : W(o) {}
};
