// cf_xlcall.h -- the slice of the Excel C API data model the wrappers of xlExport.cpp use, declared
// portably so that the x... entry points compile, load and can be tested on Linux.
//
// The reference marshals through the Excel SDK's xlcall.h (Windows only): FP12 (xlcall.h:108-113) for
// numeric ranges and XLOPER12 for strings / mixed ranges / errors.  The layouts below are the SDK's
// documented binary layouts (XCHAR = 16-bit code unit, counted strings with the length in element 0,
// INT32 rows / columns, 32-byte XLOPER12 on a 64-bit build), so an add-in shell that forwards Excel's
// pointers to these functions needs no conversion.  Only the types the wrappers read or return are
// given a member: num, str, err, multi.
#pragma once

#include <cstdint>

typedef char16_t XCHAR;
typedef int32_t  RW;
typedef int32_t  COL;

typedef struct _FP12 {
    int32_t rows;
    int32_t columns;
    double  array[1];        /* actually array[rows][columns], row major */
} FP12;

typedef struct xloper12 {
    union {
        double num;                                           /* xltypeNum */
        XCHAR* str;                                           /* xltypeStr: str[0] = length, then the characters */
        int32_t xbool;                                        /* xltypeBool */
        int32_t err;                                          /* xltypeErr */
        struct { struct xloper12* lparray; RW rows; COL columns; } array;   /* xltypeMulti, row major */
        unsigned char pad[24];                                /* the SDK's union is 24 bytes (references) */
    } val;
    uint32_t xltype;
} XLOPER12, *LPXLOPER12;

static_assert(sizeof(XLOPER12) == 32, "XLOPER12 is 32 bytes in the 64-bit Excel SDK");

#define xltypeNum     0x0001
#define xltypeStr     0x0002
#define xltypeBool    0x0004
#define xltypeErr     0x0010
#define xltypeMulti   0x0040
#define xltypeMissing 0x0080
#define xltypeNil     0x0100

#define xlerrNull  0
#define xlerrDiv0  7
#define xlerrValue 15
#define xlerrRef   23
#define xlerrName  29
#define xlerrNum   36
#define xlerrNA    42
