// archive <file> -> <file>.compressed — same command line and output name as the reference's
// `archive` (/root/reference/Compressor.cu:315-321, :427-429, Makefile:5-6).  All work is done
// by libhuffb200 on GPU 0; there is no CPU path.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../../include/huffman_b200.h"

int main(int argc, char *argv[])
{
    if (argc != 2) {                                    // C:317-321: message, exit code 0
        printf("Must provide a single file name.\n");
        return 0;
    }
    const auto t0 = std::chrono::steady_clock::now();
    hf_ctx *ctx = nullptr;
    if (hf_ctx_create(&ctx, 0, nullptr) != HF_OK) {
        fprintf(stderr, "archive: no usable B200 (sm_100) GPU\n");
        return 2;
    }
    const auto t1 = std::chrono::steady_clock::now();
    int rc = hf_archive_file(ctx, argv[1]);
    const auto t2 = std::chrono::steady_clock::now();
    if (rc != HF_OK) fprintf(stderr, "archive: error %d: %s\n", rc, hf_last_error(ctx));
    hf_ctx_destroy(ctx);
    if (const char *e = getenv("HF_TIMING"); e && e[0] && e[0] != '0') {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[hf timing] CUDA start-up + context %.1f ms, hf_archive_file %.1f ms, teardown %.1f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, t3));
    }
    return rc == HF_OK ? 0 : 2;
}
