// Stand-in for FreeImage (Windows import library only in the reference checkout): just enough for P/Texture.cpp to
// compile. The loader functions report "unknown format", so Texture's constructor leaves the object empty; the shim then
// fills Texture's members (width_, height_, scan_width_, pixel_size_, data_) from an RbTexture, which is exactly what the
// constructor would have stored after FreeImage_ConvertToRawBits.
#pragma once
typedef int FREE_IMAGE_FORMAT;
#define FIF_UNKNOWN (-1)
#define FI_RGBA_RED_MASK 0x00FF0000
#define FI_RGBA_GREEN_MASK 0x0000FF00
#define FI_RGBA_BLUE_MASK 0x000000FF
#ifndef TRUE
#define TRUE 1
#endif
inline FREE_IMAGE_FORMAT FreeImage_GetFileType(const char*, int) { return FIF_UNKNOWN; }
inline FREE_IMAGE_FORMAT FreeImage_GetFIFFromFilename(const char*) { return FIF_UNKNOWN; }
inline int FreeImage_FIFSupportsReading(FREE_IMAGE_FORMAT) { return 0; }
inline FIBITMAP* FreeImage_Load(FREE_IMAGE_FORMAT, const char*) { return nullptr; }
inline BYTE* FreeImage_GetBits(FIBITMAP*) { return nullptr; }
inline unsigned FreeImage_GetWidth(FIBITMAP*) { return 0; }
inline unsigned FreeImage_GetHeight(FIBITMAP*) { return 0; }
inline unsigned FreeImage_GetPitch(FIBITMAP*) { return 0; }
inline unsigned FreeImage_GetBPP(FIBITMAP*) { return 0; }
inline void FreeImage_ConvertToRawBits(BYTE*, FIBITMAP*, int, unsigned, unsigned, unsigned, unsigned, int) {}
inline void FreeImage_Unload(FIBITMAP*) {}
