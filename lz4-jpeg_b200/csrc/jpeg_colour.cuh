// lz4-jpeg_b200/csrc/jpeg_colour.cuh — RGB -> Y / Cr / Cb of the reference (JPEG.c:114-185), shared by the encoder
// and by the decoder (which needs it for the groups the reference leaves unprocessed).
// Included INSIDE the including file's namespace.
//
// The reference evaluates 0.299*r + 0.587*g + 0.114*b (etc.) in double and truncates.  The exact rational value
// is S/1000 with S an integer; unless S is a multiple of 1000 it is >= 1e-3 away from every integer while the
// double evaluation errs by < 1e-12, so floor(S/1000) is the answer; for S % 1000 == 0 the double expression
// itself is evaluated with explicit round-to-nearest mul/add in the reference's left-to-right order (no FMA).
__device__ __noinline__ int luma_exact(int r, int g, int b)
{
    double y = __dadd_rn(__dadd_rn(__dmul_rn(0.299, (double)r), __dmul_rn(0.587, (double)g)), __dmul_rn(0.114, (double)b));
    return (int)y & 0xFF; // implicit double -> uint8_t conversion of a value in [0, 255]
}
__device__ __forceinline__ int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
__device__ __noinline__ int cr_exact(int r, int g, int b)
{
    double v = __dadd_rn(__dsub_rn(__dsub_rn(__dmul_rn(0.439, (double)r), __dmul_rn(0.368, (double)g)), __dmul_rn(0.071, (double)b)), 128.0);
    return clamp255((int)v);
}
__device__ __noinline__ int cb_exact(int r, int g, int b)
{
    double v = __dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(-0.148, (double)r), __dmul_rn(0.291, (double)g)), __dmul_rn(0.439, (double)b)), 128.0);
    return clamp255((int)v);
}
// floor(s / 1000) for 0 <= s < 2^18, and whether s is a multiple of 1000
__device__ __forceinline__ int div1000(int s, bool &tie)
{
    const int q = (int)__umulhi((unsigned)s, 4294968u); // ceil(2^32 / 1000): exact for s < 2^22
    tie = (s - q * 1000) == 0;
    return q;
}
__device__ __forceinline__ int luma_of(int r, int g, int b)
{
    bool tie;
    const int q = div1000(299 * r + 587 * g + 114 * b, tie);
    return tie ? luma_exact(r, g, b) : q;
}
__device__ __forceinline__ int cr_of(int r, int g, int b)
{
    bool tie; // 439r - 368g - 71b + 128000 lies in [16055, 239945]: positive, so (int) truncation is floor and clamp is idle
    const int q = div1000(439 * r - 368 * g - 71 * b + 128000, tie);
    return tie ? cr_exact(r, g, b) : q;
}
__device__ __forceinline__ int cb_of(int r, int g, int b)
{
    bool tie;
    const int q = div1000(-148 * r - 291 * g + 439 * b + 128000, tie);
    return tie ? cb_exact(r, g, b) : q;
}

