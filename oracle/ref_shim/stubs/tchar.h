/* stub of the MSVC header pulled in by the reference's PSBA/stdafx.h: maps the three
 * Microsoft "secure" stdio calls the loaders use onto their ISO C equivalents. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
typedef int errno_t;
static inline errno_t fopen_s(FILE **f, const char *name, const char *mode) { *f = fopen(name, mode); return *f ? 0 : 1; }
#define fscanf_s fscanf
#define sscanf_s sscanf
