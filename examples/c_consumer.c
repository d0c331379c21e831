/* A plain-C consumer of the drop-in boundary (include/dsocr.h): what the reference's FFI (cgo-style `extern "C"` from Rust,
 * INTEGRATION.md section 2) binds.  Uses only host-side entry points, so it runs without a GPU:
 *   - writes a small DSQ snapshot with the library's writer and reads it back with the library's reader,
 *   - runs the integer preprocessing on a synthetic page,
 *   - shows that creating an engine without a CUDA device fails loudly (there is no CPU fallback).
 * Build: gcc -std=c99 -I include examples/c_consumer.c -L deepseek-ocr.rs_b200/lib -ldsocr -Wl,-rpath,$PWD/deepseek-ocr.rs_b200/lib */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "dsocr.h"

int main(int argc, char** argv) {
  const char* out = argc > 1 ? argv[1] : "c_consumer_snapshot";
  char path[1024];
  float w[2 * 32], bias[2] = {0.5f, -0.25f};
  int i, n_tiles = -1, cw = 0, ch = 0;
  dsocr_dsq_writer* wr = NULL;
  dsocr_dsq_header hdr;
  dsocr_dsq_record rec[2];
  dsocr_vision_settings vs;
  unsigned char* page;
  unsigned char* global;
  unsigned char* tiles;
  dsocr_engine* eng = NULL;

  printf("%s\n", dsocr_version());
  for (i = 0; i < 64; ++i) w[i] = (float)i * 0.25f - 3.0f;
  if (dsocr_dsq_writer_create(out, "candle-test", "unit-test", "CPU", 8, &wr) != DSOCR_OK) { printf("create: %s\n", dsocr_last_error()); return 1; }
  if (dsocr_dsq_writer_add_tensor(wr, "linear.weight", 2, 32, 8, w, bias) != DSOCR_OK) { printf("add: %s\n", dsocr_last_error()); return 1; }
  if (dsocr_dsq_writer_finalize(wr) != DSOCR_OK) { printf("finalize: %s\n", dsocr_last_error()); return 1; }
  snprintf(path, sizeof(path), "%s.dsq", out);
  if (dsocr_dsq_inspect(path, &hdr, rec, 2) != DSOCR_OK) { printf("inspect: %s\n", dsocr_last_error()); return 1; }
  printf("snapshot: %u tensor(s), `%s` [%u, %u] dtype %u, %llu payload bytes, %llu bias bytes\n", hdr.tensor_count, rec[0].name,
         rec[0].out_dim, rec[0].in_dim, rec[0].q_dtype, (unsigned long long)rec[0].q_len, (unsigned long long)rec[0].bias_len);
  if (hdr.tensor_count != 1 || rec[0].q_len != 2 * 34 || strcmp(rec[0].name, "linear.weight") != 0) return 2;

  vs.base_size = 1024; vs.image_size = 640; vs.crop_mode = 1;
  page = (unsigned char*)malloc(1654 * 2339 * 3);
  global = (unsigned char*)malloc(1024 * 1024 * 3);
  tiles = (unsigned char*)malloc((size_t)9 * 640 * 640 * 3);
  memset(page, 255, 1654 * 2339 * 3);
  if (dsocr_preprocess(page, 1654, 2339, vs, global, tiles, &n_tiles, &cw, &ch) != DSOCR_OK) { printf("preprocess: %s\n", dsocr_last_error()); return 1; }
  printf("A4 page: grid %dx%d, %d tiles, %d image tokens\n", cw, ch, n_tiles, dsocr_image_token_count(1024, 640, 1, cw, ch));
  if (cw != 2 || ch != 3 || n_tiles != 6 || dsocr_image_token_count(1024, 640, 1, cw, ch) != 903) return 3;
  free(page); free(global); free(tiles);

  if (dsocr_engine_create("missing_config.json", "missing.safetensors", NULL, 0, DSOCR_BF16, &eng) == DSOCR_OK) {
    printf("engine created\n");
    dsocr_engine_destroy(eng);
  } else {
    printf("engine_create failed as expected here: %s\n", dsocr_last_error());
  }
  return 0;
}
