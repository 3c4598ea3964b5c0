/* fixed_seed.c -- TEST INFRASTRUCTURE ONLY (oracle/).  LD_PRELOAD helper for the drop-in tests.
 *
 * codeml / baseml seed their random-number generator from /dev/urandom (SetSeed, src/tools.c:1378-1409) and perturb the
 * initial values of every fit with it (src/codeml.c:2102-2144, src/baseml.c:1242-1277), so two runs of the SAME binary on the
 * same files start at different points.  A test that compares the unmodified reference with the drop-in wants both to start
 * at the same point: with this library preloaded and PB200_TEST_SEED set, fopen("/dev/urandom") hands back a stream of bytes
 * derived from that number; every other fopen goes to libc untouched.  The programs themselves are not modified.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static FILE *fake_urandom(void)
{
   static unsigned char bytes[64];
   unsigned int s = (unsigned int)strtoul(getenv("PB200_TEST_SEED"), NULL, 10), i;
   for (i = 0; i < sizeof(bytes); i++) { s = s * 1103515245u + 12345u; bytes[i] = (unsigned char)(s >> 16); }
   bytes[3] &= 0x3f;                     /* keep the 32-bit seed positive and small enough for `seed * 2 - 1` */
   return fmemopen(bytes, sizeof(bytes), "r");
}

FILE *fopen(const char *path, const char *mode)
{
   static FILE *(*real)(const char *, const char *) = NULL;
   if (!real) real = (FILE *(*)(const char *, const char *))dlsym(RTLD_NEXT, "fopen");
   if (path && strcmp(path, "/dev/urandom") == 0 && getenv("PB200_TEST_SEED")) return fake_urandom();
   return real(path, mode);
}

FILE *fopen64(const char *path, const char *mode)
{
   static FILE *(*real)(const char *, const char *) = NULL;
   if (!real) real = (FILE *(*)(const char *, const char *))dlsym(RTLD_NEXT, "fopen64");
   if (path && strcmp(path, "/dev/urandom") == 0 && getenv("PB200_TEST_SEED")) return fake_urandom();
   return real(path, mode);
}
