// extract <file.compressed> -> ./DECOMPRESSED_FILE — same command line, messages and output
// naming as the reference's `extract` (/root/reference/Decompressor.cu:47-63, :104-105,
// :185-219), but the decode runs on GPU 0 through libhuffb200.
#include <cstdio>

#include "../../include/huffman_b200.h"

int main(int argc, char *argv[])
{
    if (argc != 2) {                                    // D:51-56: message, exit code 1
        printf("Missing compressed file name.\nUsage: './extract <compressed_file_name>'\n");
        return 1;
    }
    hf_ctx *ctx = nullptr;
    if (hf_ctx_create(&ctx, 0, nullptr) != HF_OK) {
        fprintf(stderr, "extract: no usable B200 (sm_100) GPU\n");
        return 2;
    }
    int rc = hf_extract_file(ctx, argv[1]);
    if (rc != HF_OK) fprintf(stderr, "extract: error %d: %s\n", rc, hf_last_error(ctx));
    hf_ctx_destroy(ctx);
    return rc == HF_OK ? 0 : 2;
}
