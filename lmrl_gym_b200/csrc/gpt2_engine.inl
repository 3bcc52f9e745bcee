// placeholder until the decoder lands
void lmrl_gpt2_free(lmrl_engine*) {}
