// Link-time stand-ins for the FreeImage entry points the reference's Texture.cpp
// references (Texture.cpp:25-53,104-153,167).  LoadedTexture is never constructed
// on the hot path (no BASELINE config uses an image texture), so none of these is
// ever called; they exist only so oracle/_ref links without the vendored 400 kLoC
// FreeImage tree.  TEST INFRASTRUCTURE, not product code.
#include <cstdlib>
#include "FreeImage.h"
extern "C" {
void DLL_CALLCONV FreeImage_Initialise(BOOL) {}
void DLL_CALLCONV FreeImage_DeInitialise() {}
FREE_IMAGE_FORMAT DLL_CALLCONV FreeImage_GetFileType(const char*, int) { return FIF_UNKNOWN; }
FIBITMAP* DLL_CALLCONV FreeImage_Load(FREE_IMAGE_FORMAT, const char*, int) { abort(); }
unsigned DLL_CALLCONV FreeImage_GetWidth(FIBITMAP*) { abort(); }
unsigned DLL_CALLCONV FreeImage_GetHeight(FIBITMAP*) { abort(); }
unsigned DLL_CALLCONV FreeImage_GetBPP(FIBITMAP*) { abort(); }
FREE_IMAGE_TYPE DLL_CALLCONV FreeImage_GetImageType(FIBITMAP*) { abort(); }
FIBITMAP* DLL_CALLCONV FreeImage_AllocateT(FREE_IMAGE_TYPE, int, int, int, unsigned, unsigned, unsigned) { abort(); }
BOOL DLL_CALLCONV FreeImage_SetPixelColor(FIBITMAP*, unsigned, unsigned, RGBQUAD*) { abort(); }
BOOL DLL_CALLCONV FreeImage_GetPixelColor(FIBITMAP*, unsigned, unsigned, RGBQUAD*) { abort(); }
BYTE* DLL_CALLCONV FreeImage_GetScanLine(FIBITMAP*, int) { abort(); }
FIBITMAP* DLL_CALLCONV FreeImage_Copy(FIBITMAP*, int, int, int, int) { abort(); }
FIBITMAP* DLL_CALLCONV FreeImage_Rescale(FIBITMAP*, int, int, FREE_IMAGE_FILTER) { abort(); }
}
