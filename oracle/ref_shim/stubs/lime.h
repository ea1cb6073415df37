// TEST INFRASTRUCTURE stand-in (see quda.h in this directory): the c-lime declarations the reference headers mention
#pragma once
#include <stdio.h>
#include <stdint.h>
typedef uint64_t n_uint64_t;
typedef struct LimeReader_s LimeReader;
typedef struct LimeWriter_s LimeWriter;
typedef struct LimeRecordHeader_s LimeRecordHeader;
#define LIME_EOF (-4)
LimeReader *limeCreateReader(FILE *fp);
void limeDestroyReader(LimeReader *r);
int limeReaderNextRecord(LimeReader *r);
char *limeReaderType(LimeReader *r);
n_uint64_t limeReaderBytes(LimeReader *r);
int limeReaderReadData(void *dest, n_uint64_t *nbytes, LimeReader *r);
