/* aux oracle units (PosMLP / CDF / SH) are added in mb_oracle_aux.c */
