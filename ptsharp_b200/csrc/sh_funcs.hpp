// Real spherical harmonics Y_l^m, l <= 4, as SphericalHarmonic.shFunc selects them (SH.cs:106-338).  The reference writes the
// coefficients as float literals that are widened to double and multiplies by the double-typed accessors of a (float-stored)
// unit Vector, left to right; this header restates each product in that order.  Shared by the host (marching cubes over the
// harmonic solid, host/mc.cpp) and the device (NormalAt / MaterialAt of a hit, pt_device.cuh).
#pragma once
#if defined(__CUDACC__)
#define PT_SH_HD __host__ __device__ __forceinline__
#else
#define PT_SH_HD inline
#endif

// d = a unit direction with float-valued components (Vector.Normalize()).  Returns NaN for an unsupported (l, m) - the reference
// prints "unsupported spherical harmonic" and then dereferences a null delegate.
PT_SH_HD double sh_eval(int l, int m, double x, double y, double z) {
    const double F2 = (double)2.0f, F3 = (double)3.0f, F4 = (double)4.0f, F7 = (double)7.0f, F1 = (double)1.0f;
    switch (l * 16 + (m + 4)) {
        case 0 * 16 + 4: return (double)0.282095f;                                                             // sh00
        case 1 * 16 + 3: return (double)-0.488603f * y;                                                        // sh1n1
        case 1 * 16 + 4: return (double)0.488603f * z;                                                         // sh10
        case 1 * 16 + 5: return (double)-0.488603f * x;                                                        // sh1p1
        case 2 * 16 + 2: return (double)1.092548f * x * y;                                                     // sh2n2
        case 2 * 16 + 3: return (double)-1.092548f * y * z;                                                    // sh2n1
        case 2 * 16 + 4: return (double)0.315392f * (-x * x - y * y + F2 * z * z);                             // sh20
        case 2 * 16 + 5: return (double)-1.092548f * x * z;                                                    // sh2p1
        case 2 * 16 + 6: return (double)0.546274f * (x * x - y * y);                                           // sh2p2
        case 3 * 16 + 1: return (double)-0.590044f * y * (F3 * x * x - y * y);                                 // sh3n3
        case 3 * 16 + 2: return (double)2.890611f * x * y * z;                                                 // sh3n2
        case 3 * 16 + 3: return (double)-0.457046f * y * (F4 * z * z - x * x - y * y);                         // sh3n1
        case 3 * 16 + 4: return (double)0.373176f * z * (F2 * z * z - F3 * x * x - F3 * y * y);                // sh30
        case 3 * 16 + 5: return (double)-0.457046f * x * (F4 * z * z - x * x - y * y);                         // sh3p1
        case 3 * 16 + 6: return (double)1.445306f * z * (x * x - y * y);                                       // sh3p2
        case 3 * 16 + 7: return (double)-0.590044f * x * (x * x - F3 * y * y);                                 // sh3p3
        case 4 * 16 + 0: return (double)2.503343f * x * y * (x * x - y * y);                                   // sh4n4
        case 4 * 16 + 1: return (double)-1.770131f * y * z * (F3 * x * x - y * y);                             // sh4n3
        case 4 * 16 + 2: return (double)0.946175f * x * y * (F7 * z * z - F1);                                 // sh4n2
        case 4 * 16 + 3: return (double)-0.669047f * y * z * (F7 * z * z - F3);                                // sh4n1
        case 4 * 16 + 4: { const double z2 = z * z; return (double)0.105786f * ((double)35.0f * z2 * z2 - (double)30.0f * z2 + F3); }  // sh40
        case 4 * 16 + 5: return (double)-0.669047f * x * z * (F7 * z * z - F3);                                // sh4p1
        case 4 * 16 + 6: return (double)0.473087f * (x * x - y * y) * (F7 * z * z - F1);                       // sh4p2
        case 4 * 16 + 7: return (double)-1.770131f * x * z * (x * x - F3 * y * y);                             // sh4p3
        case 4 * 16 + 8: { const double x2 = x * x, y2 = y * y; return (double)0.625836f * (x2 * (x2 - F3 * y2) - y2 * (F3 * x2 - y2)); }  // sh4p4
        default: return x - x + (y - y) / (z - z);  // NaN (unsupported pairs are rejected before they get here)
    }
}
PT_SH_HD bool sh_supported(int l, int m) { return l >= 0 && l <= 4 && m >= -l && m <= l; }
