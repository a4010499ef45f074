// archive <file> -> <file>.compressed — same command line and output name as the reference's
// `archive` (/root/reference/Compressor.cu:315-321, :427-429, Makefile:5-6).  All work is done
// by libhuffb200 on GPU 0; there is no CPU path.
#include <cstdio>

#include "../../include/huffman_b200.h"

int main(int argc, char *argv[])
{
    if (argc != 2) {                                    // C:317-321: message, exit code 0
        printf("Must provide a single file name.\n");
        return 0;
    }
    hf_ctx *ctx = nullptr;
    if (hf_ctx_create(&ctx, 0, nullptr) != HF_OK) {
        fprintf(stderr, "archive: no usable B200 (sm_100) GPU\n");
        return 2;
    }
    int rc = hf_archive_file(ctx, argv[1]);
    if (rc != HF_OK) fprintf(stderr, "archive: error %d: %s\n", rc, hf_last_error(ctx));
    hf_ctx_destroy(ctx);
    return rc == HF_OK ? 0 : 2;
}
