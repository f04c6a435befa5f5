/* Pure C99 consumer of include/onb.h: proves the header is C-clean and the library links from C. Calls only the host-side
 * helpers (no device needed) and then checks that onb_create fails loudly without a GPU / works with one. */
#include <stdio.h>
#include <string.h>
#include "onb.h"

int main(void) {
    uint8_t deck[5] = {1, 2, 0, 3, 11};
    onb_state s;
    uint32_t att[800];
    uint8_t d[5];
    onb_config cfg;
    onb_ctx* ctx = NULL;
    int32_t rc;
    if (onb_version() != ONB_VERSION) return 1;
    if (onb_start_states(deck, 1, &s) != ONB_OK) return 2;
    if (s.pawns[0] != 0x00000D80u || s.kings[1] != 0x20000000u || s.side != ONB_RED) return 3; /* Horse carries the Red stamp */
    if (onb_attack_maps(att) != ONB_OK) return 4;
    if (att[(0 * 16 + 0) * 25 + 22] != ((1u << (31 - 12)))) return 5; /* Red Tiger from c1: only c3 (two rows up) is on the board... */
    if (onb_deal(42, 7, 0, d) != ONB_OK) return 6;
    printf("deal %d %d %d %d %d  rng %u  action %u\n", d[0], d[1], d[2], d[3], d[4], onb_rand_u32(1, 2, 3, 4),
           (unsigned)ONB_ACTION(1, 24, 18, ONB_PAWN));
    memset(&cfg, 0, sizeof cfg);
    cfg.n_games = 4;
    rc = onb_create(&cfg, &ctx);
    if (rc == ONB_OK) {
        onb_state out[4];
        if (onb_env_reset(ctx, deck, 1, 0) != ONB_OK) return 7;
        if (onb_env_get_states(ctx, out, 0, 4) != ONB_OK) return 8;
        if (memcmp(&out[3], &s, sizeof s) != 0) return 9;
        onb_destroy(ctx);
        printf("C ABI OK (device)\n");
    } else {
        if (rc != ONB_E_CUDA || strstr(onb_last_error(NULL), "no CPU fallback") == NULL) return 10;
        printf("C ABI OK (no device: %s)\n", onb_last_error(NULL));
    }
    return 0;
}
