/* loglik_c99.c — the C-ABI (include/dmpc.h) from plain C99, the way a foreign-function binding sees it: no C++, no
 * Python, no torch.  One logLikCpp-style create on a four-tip tree, one proposal of new between-cluster matrices, its
 * rejection, and the release of the handle — the call sequence of the reference's R driver
 * (/root/reference/R/DMphyClusSource.R:216-255) in miniature.
 *
 *   gcc -std=c99 -Wall -Wextra -pedantic -Iinclude examples/loglik_c99.c -Ldmphyclus_b200 -ldmpc \
 *       -Wl,-rpath,$PWD/dmphyclus_b200 -o /tmp/loglik_c99 && /tmp/loglik_c99
 *
 * Exit status: 0 = evaluated on the GPU; 3 = the library refused because no sm_100-class device is usable (there is no
 * CPU fallback: tests/test_c_consumer.py expects exactly this on a box without a GPU); 1 = any other failure. */
#include <stdio.h>
#include <string.h>

#include "dmpc.h"

static int failed(const char* what, int rc) {
  fprintf(stderr, "%s: status %d: %s\n", what, rc, dmpc_last_error());
  return rc == DMPC_ERR_CUDA ? 3 : 1;
}

int main(void) {
  /* ((1,2),(3,4)) with root 5: phylo$edge, column-major, 1-based (parents, then children), rows in pre-order */
  const int32_t edge[12] = {5, 6, 6, 5, 7, 7, /* children */ 6, 1, 2, 7, 3, 4};
  const double mrcas[2] = {6, 7};                       /* two clusters: {1,2} and {3,4} */
  const double pi[4] = {0.39, 0.22, 0.17, 0.22};        /* a, t, c, g */
  enum { R = 2, TIPS = 4, LOCI = 8 };
  double within[R * 16], between[R * 16], between2[R * 16];
  const char* rows[TIPS] = {"atcgatcg", "atcgaccg", "ttcgrtcn", "ttcg-tca"};
  char cells[TIPS * LOCI];
  uint8_t masks[TIPS * LOCI];
  dmpc_options opt;
  dmpc_tree* tree = NULL;
  double ll = 0.0, ll2 = 0.0, ll3 = 0.0;
  int r, i, j, rc;

  /* row-stochastic 4x4 matrices, column-major as R hands them over: P(i,j) at [i + 4 j] */
  for (r = 0; r < R; r++)
    for (i = 0; i < 4; i++)
      for (j = 0; j < 4; j++) {
        const double stay_w = 0.97 - 0.02 * r, stay_b = 0.85 - 0.05 * r, stay_b2 = 0.80 - 0.05 * r;
        within[16 * r + i + 4 * j] = i == j ? stay_w : (1.0 - stay_w) / 3.0;
        between[16 * r + i + 4 * j] = i == j ? stay_b : (1.0 - stay_b) / 3.0;
        between2[16 * r + i + 4 * j] = i == j ? stay_b2 : (1.0 - stay_b2) / 3.0;
      }
  for (i = 0; i < TIPS; i++) memcpy(cells + i * LOCI, rows[i], LOCI);
  if (dmpc_convert_alignment(cells, TIPS * LOCI, "atcg", masks) != 0) return failed("dmpc_convert_alignment", DMPC_ERR_ARG);
  printf("%s\nmasks of tip 3: ", dmpc_version());
  for (j = 0; j < LOCI; j++) printf("%x", (unsigned)masks[2 * LOCI + j]);
  printf("\n");

  dmpc_default_options(&opt);
  rc = dmpc_loglik_create(edge, 6, mrcas, 2, pi, within, between, R, masks, TIPS, LOCI, 1, 1, &opt, &tree, &ll);
  if (rc != DMPC_OK) return failed("dmpc_loglik_create", rc);
  rc = dmpc_new_between_transprobs_loglik(tree, within, between2, 2, pi, &ll2);          /* propose */
  if (rc != DMPC_OK) return failed("dmpc_new_between_transprobs_loglik", rc);
  rc = dmpc_restore_previous_config(tree, NULL, 0, 1, 1, 0, NULL, 0, 0);                 /* reject  */
  if (rc != DMPC_OK) return failed("dmpc_restore_previous_config", rc);
  rc = dmpc_recompute_all(tree, within, between, pi, &ll3);                              /* the committed state again */
  if (rc != DMPC_OK) return failed("dmpc_recompute_all", rc);
  rc = dmpc_final_deallocate(tree);
  if (rc != DMPC_OK) return failed("dmpc_final_deallocate", rc);
  printf("logLik %.17g  proposal %.17g  after the rejection %.17g\n", ll, ll2, ll3);
  return ll3 == ll && ll2 != ll ? 0 : 1;
}
