/* TEST INFRASTRUCTURE — stand-in for the tokenizers-cpp C ABI (reference
 * tokenizers-cpp/include/tokenizers_c.h:17-50), which is a Rust static library that cannot be
 * built in this image (no cargo/rustc).  Tokenisation is OUT OF SCOPE for the B200 engine
 * ("tokenizers-cpp stays on the host"); this stub only exists so the
 * UNCHANGED reference main.c / src/tokenizer.c can be linked and driven end to end against
 * libgliclass_b200.so.  It is deterministic and mirrored in tests/ref_driver.py:
 *   [CLS]=1, then per whitespace-separated word: FNV-1a(word) mapped into [3, class_token),
 *   "<<LABEL>>" -> class_token, "<<SEP>>" -> sep_token (both split words), then [SEP]=2.
 * The "tokenizer.json" it is created from is {"class_token": N, "sep_token": M}.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tokenizers_c.h"

typedef struct { int class_token, sep_token; } StubTok;

static int find_int(const char* json, const char* key, int dflt) {
  const char* p = strstr(json, key);
  if (!p) return dflt;
  p = strchr(p, ':');
  if (!p) return dflt;
  return atoi(p + 1);
}

TokenizerHandle tokenizers_new_from_str(const char* json, size_t len) {
  (void)len;
  StubTok* t = (StubTok*)malloc(sizeof(StubTok));
  t->class_token = find_int(json, "\"class_token\"", 128001);
  t->sep_token = find_int(json, "\"sep_token\"", 128002);
  return t;
}

static int word_id(const StubTok* t, const char* s, size_t n) {
  uint32_t h = 2166136261u;
  for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)s[i]; h *= 16777619u; }
  return 3 + (int)(h % (uint32_t)(t->class_token - 3));
}

static void encode_one(const StubTok* t, const char* s, size_t n, int special, TokenizerEncodeResult* r) {
  int* ids = (int*)malloc((n + 2) * sizeof(int));
  size_t k = 0, i = 0;
  if (special) ids[k++] = 1;
  while (i < n) {
    if (s[i] == ' ' || s[i] == '\t' || s[i] == '\n') { ++i; continue; }
    if (n - i >= 9 && !strncmp(s + i, "<<LABEL>>", 9)) { ids[k++] = t->class_token; i += 9; continue; }
    if (n - i >= 7 && !strncmp(s + i, "<<SEP>>", 7)) { ids[k++] = t->sep_token; i += 7; continue; }
    size_t j = i;
    while (j < n && s[j] != ' ' && s[j] != '\t' && s[j] != '\n' && strncmp(s + j, "<<LABEL>>", 9) && strncmp(s + j, "<<SEP>>", 7)) ++j;
    ids[k++] = word_id(t, s + i, j - i);
    i = j;
  }
  if (special) ids[k++] = 2;
  r->token_ids = ids;
  r->len = k;
}

void tokenizers_encode_batch(TokenizerHandle h, const char** data, size_t* len, size_t num, int special,
                             TokenizerEncodeResult* results) {
  for (size_t i = 0; i < num; ++i) encode_one((const StubTok*)h, data[i], len[i], special, &results[i]);
}

void tokenizers_free_encode_results(TokenizerEncodeResult* results, size_t num) {
  for (size_t i = 0; i < num; ++i) free(results[i].token_ids);
  /* like the Rust side (tokenizers-cpp/rust/src/lib.rs:189-196) only the id arrays are freed */
}

void tokenizers_free(TokenizerHandle h) { free(h); }
