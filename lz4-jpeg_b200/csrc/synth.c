/* lz4-jpeg_b200/csrc/synth.c — workload generators of the reference's harnesses (host code).
 *
 * ljb_synth_text : Experiment/random_extract.c:8-71 — copy `passage` bytes of the corpus starting at
 *                  rng() % (corpus_len - passage) (:36), CR/LF -> space (:49-53), repeated until n bytes.
 * ljb_synth_image: Experiment/random_image.c:58-77 — r,g,b = rng() % 256 each, a = 255.
 * The reference seeds with time(NULL) / not at all; a splitmix64 stream with an explicit seed is used
 * so that every run (CPU baseline and GPU) sees the same bytes. */
#include <stddef.h>
#include <stdint.h>

static inline uint64_t splitmix64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void ljb_synth_text(const uint8_t *corpus, size_t corpus_len, uint64_t seed, size_t passage, uint8_t *out, size_t n)
{
    uint64_t s = seed;
    size_t o = 0;
    if (!corpus || !out || passage == 0 || corpus_len <= passage) return;
    while (o < n) {
        size_t start = (size_t)(splitmix64(&s) % (uint64_t)(corpus_len - passage));
        size_t take = (n - o < passage) ? n - o : passage;
        for (size_t k = 0; k < take; ++k) {
            uint8_t c = corpus[start + k];
            out[o + k] = (c == '\n' || c == '\r') ? (uint8_t)' ' : c;
        }
        o += take;
    }
}

void ljb_synth_image(uint64_t seed, int w, int h, uint8_t *rgba)
{
    uint64_t s = seed;
    size_t npx = (size_t)w * (size_t)h;
    for (size_t i = 0; i < npx; i += 2) { /* one 64-bit draw feeds the six colour bytes of two pixels */
        uint64_t z = splitmix64(&s);
        rgba[4 * i + 0] = (uint8_t)(z);
        rgba[4 * i + 1] = (uint8_t)(z >> 8);
        rgba[4 * i + 2] = (uint8_t)(z >> 16);
        rgba[4 * i + 3] = 255;
        if (i + 1 < npx) {
            rgba[4 * i + 4] = (uint8_t)(z >> 24);
            rgba[4 * i + 5] = (uint8_t)(z >> 32);
            rgba[4 * i + 6] = (uint8_t)(z >> 40);
            rgba[4 * i + 7] = 255;
        }
    }
}
