/*
 * cr_main.c -- host side of the drop-in `comprolz` / `comprop` command lines.
 *
 * Keeps the reference's CLI (src/rolzmain/main.c:67-112, src/ropmain/main.c) and container format
 * (src/main.c:67-79,90-94,153-206), but instead of looping over blocks and calling
 * filter_inplace / dictionary_encode / lzencode per block (src/main.c:174-206) it hands the whole input to
 * crgpu_compress() -- adaptive models carry across blocks, so the GPU needs the whole chain.
 * The library is loaded with dlopen so that a missing CUDA build fails loudly at run time; there is no CPU path.
 */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <libgen.h>
#include <unistd.h>
#include "../../include/crgpu.h"

#ifndef CR_VARIANT
#define CR_VARIANT 0
#endif
static const char* NAME = CR_VARIANT == 0 ? "comprolz" : CR_VARIANT == 1 ? "comprop" : "comprox";

static uint32_t cr_split_size = 16 * 1048576;    /* src/main.c:62 */
static int cr_filt_enable = 0, cr_prec_enable = 0, flexible_parsing = 0, quiet = 0;
static int match_limit = 0;                      /* comprox -m (src/roxmain/main.c:94-98); 0 = default 40 */

static void usage(void) {
    fprintf(stderr,
            "to compress:   %s [SWITCH] e [input] [output]\n"
            "to decompress: %s          d [input] [output]\n"
            "work with standard I/O streams if filenames are not given.\n\n"
            "optional SWITCH:\n"
            "   -b  set block size(MB), default = 16.\n"
            "   -p  work as a precompressor.\n"
            "   -F  use PE/ELF/BMP filter.\n"
            "%s%s"
            "   -q  quiet mode.\n", NAME, NAME, CR_VARIANT != 1 ? "   -f  use flexible parsing.\n" : "",
            CR_VARIANT == 2 ? "   -m  set maximum searching depth for LZ77 matching, default = 40.\n" : "");
}

/* src/rolzmain/main.c:67-112 */
static int process_arguments(int argc, char** argv) {
    while (argc >= 2 && argv[1][0] == '-') {
        switch (argv[1][1]) {
            case 'b': if ((cr_split_size = (uint32_t)atoi(argv[1] + 2) * 1048576u) == 0) goto bad; break;
            case 'p': if (argv[1][2]) goto bad; cr_prec_enable = 1; break;
            case 'F': if (argv[1][2]) goto bad; cr_filt_enable = 1; break;
            case 'q': if (argv[1][2]) goto bad; quiet = 1; break;
            case 'f': if (CR_VARIANT == 1 || argv[1][2]) goto bad; flexible_parsing = 1; break;
            case 'm': if (CR_VARIANT != 2 || (match_limit = atoi(argv[1] + 2)) <= 0) goto bad; break;
            default:
            bad:
                fprintf(stderr, "invalid switch '%s'.\n", argv[1]);
                return 0;
        }
        memmove(argv + 1, argv + 2, (size_t)(argc - 2) * sizeof(char*));
        argc--;
    }
    return argc;
}

static uint8_t* read_all(FILE* f, uint64_t* n) {
    size_t cap = 1 << 20, len = 0;
    uint8_t* p = malloc(cap);
    for (;;) {
        size_t r = fread(p + len, 1, cap - len, f);
        len += r;
        if (r == 0) break;
        if (len == cap) { cap *= 2; p = realloc(p, cap); }
    }
    *n = len;
    return p;
}

int main(int argc, char** argv) {
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    if ((argc = process_arguments(argc, argv)) == 0) return -1;
    if (!(argc >= 2 && argc <= 4 && (!strcmp(argv[1], "e") || !strcmp(argv[1], "d")))) { usage(); return -1; }
    const int decode = !strcmp(argv[1], "d");
    FILE* src = argc >= 3 ? fopen(argv[2], "rb") : stdin;
    FILE* dst = argc >= 4 ? fopen(argv[3], "wb") : stdout;
    if (!src || !dst) { perror("fopen()"); return -1; }

    /* libcrgpu.so sits next to the package (../comprox_b200/) or is named by CRGPU_LIB */
    char path[4096];
    const char* env = getenv("CRGPU_LIB");
    if (env) snprintf(path, sizeof path, "%s", env);
    else {
        char exe[4096]; ssize_t k = readlink("/proc/self/exe", exe, sizeof exe - 1); exe[k > 0 ? k : 0] = 0;
        snprintf(path, sizeof path, "%s/../comprox_b200/libcrgpu.so", dirname(exe));
    }
    void* lib = dlopen(path, RTLD_NOW);
    if (!lib) { fprintf(stderr, "%s: cannot load %s: %s (no CPU fallback)\n", NAME, path, dlerror()); return -1; }
    int (*p_create)(crgpu_handle**, int, int, void*) = dlsym(lib, "crgpu_create");
    int (*p_compress)(crgpu_handle*, const crgpu_config*, const uint8_t*, uint64_t, uint8_t*, uint64_t, uint64_t*) = dlsym(lib, "crgpu_compress");
    int (*p_decompress)(crgpu_handle*, const uint8_t*, uint64_t, uint8_t*, uint64_t, uint64_t*) = dlsym(lib, "crgpu_decompress");
    uint64_t (*p_bound)(uint64_t, uint32_t) = dlsym(lib, "crgpu_compress_bound");
    const char* (*p_err)(int) = dlsym(lib, "crgpu_strerror");
    int (*p_option)(crgpu_handle*, const char*, int64_t) = dlsym(lib, "crgpu_set_option");
    void (*p_destroy)(crgpu_handle*) = dlsym(lib, "crgpu_destroy");

    uint64_t n = 0, out_n = 0;
    uint8_t* in = read_all(src, &n);
    if (!quiet && !decode) fprintf(stderr, "compressing %s to %s, block_size = %uMB...\n", argc >= 3 ? argv[2] : "<stdin>", argc >= 4 ? argv[3] : "<stdout>", cr_split_size / 1048576);
    if (!quiet && decode) fprintf(stderr, "decompressing %s to %s...\n", argc >= 3 ? argv[2] : "<stdin>", argc >= 4 ? argv[3] : "<stdout>");
    crgpu_handle* h = NULL;
    int rc = p_create(&h, CR_VARIANT, getenv("CRGPU_DEVICE") ? atoi(getenv("CRGPU_DEVICE")) : 0, NULL);
    if (rc) { fprintf(stderr, "%s: %s\n", NAME, p_err(rc)); return -1; }
    if (match_limit && (rc = p_option(h, "match_limit", match_limit)) != 0) { fprintf(stderr, "%s: %s\n", NAME, p_err(rc)); return -1; }
    crgpu_config cfg = { cr_split_size, cr_filt_enable, cr_prec_enable, flexible_parsing, 0 };
    uint64_t cap = decode ? 64 : p_bound(n, cr_split_size);
    uint8_t* out = malloc(cap);
    if (decode) {                                   /* the container does not store the raw size: grow until it fits */
        for (cap = n * 4 + (1u << 20);; ) {
            out = realloc(out, cap);
            out_n = 0;
            rc = p_decompress(h, in, n, out, cap, &out_n);
            if (rc != CRGPU_ERR_ARG || out_n <= cap) break;        /* out_n > cap: the size needed; the second call resumes, it does not decode twice */
            cap = out_n;
        }
    } else
    rc = p_compress(h, &cfg, in, n, out, cap, &out_n);
    if (rc) { fprintf(stderr, "%s: %s\n", NAME, p_err(rc)); return -1; }
    if (fwrite(out, 1, out_n, dst) != out_n) { perror("fwrite()"); return -1; }
    fclose(dst);
    p_destroy(h);
    gettimeofday(&t1, NULL);
    if (!quiet) {
        double s = (double)(t1.tv_sec - t0.tv_sec) + (t1.tv_usec - t0.tv_usec) / 1e6;
        fprintf(stderr, "%llu bytes => %llu bytes\n\nencode-speed:   %.3lf MB/s\ncost-time:      %.3lf s\ncompress-ratio: %.3lf\n",
                (unsigned long long)n, (unsigned long long)out_n, n / 1048576.0 / s, s, n ? (double)out_n / n : 0.0);
    }
    free(in); free(out);
    return 0;
}
