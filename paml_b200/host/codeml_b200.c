/* codeml_b200.c -- codeml with its likelihood function on the GPU: the reference program, textually included and unmodified,
   inside the wrapper of wrap_b200.inc (which explains the interposition). */
#define PB200_PROGRAM "codeml.c"
#include "wrap_b200.inc"
