/* kb_post_oracle.c -- placeholder for the post-mapping oracles (protein Gotoh,
 * extract/translate); filled in when those SURVEY.md section 8(f) rows are built.
 * TEST INFRASTRUCTURE ONLY. */
int kbo_post_oracle_version(void) { return 0; }
