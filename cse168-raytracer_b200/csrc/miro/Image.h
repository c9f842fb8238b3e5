// Image.h -- 8-bit RGB framebuffer of the host API layer (reference Image.h / Image.cpp): Map() truncation
// (Image.cpp:47-52), bottom-up rows, flipped binary PPM writer (Image.cpp:98-115).
#ifndef MIROHOST_IMAGE_H
#define MIROHOST_IMAGE_H
#include "Vector3.h"

class Image {
public:
    struct Pixel {
        unsigned char r, g, b;
        Pixel() : r(0), g(0), b(0) {}
        Pixel(unsigned char ir, unsigned char ig, unsigned char ib) : r(ir), g(ig), b(ib) {}
    };
    Image() : m_pixels(0), m_width(0), m_height(0), m_pinned(false) {}
    ~Image() { release(); }
    void resize(int width, int height);
    void setPixel(int x, int y, const Vector3& p);
    void setPixel(int x, int y, const Pixel& p);
    void clear(const Vector3& c);
    void writePPM(const char* pcFile);
    unsigned char* getCharPixels() { return (unsigned char*)m_pixels; }
    int width() const { return m_width; }
    int height() const { return m_height; }
private:
    void release();
    Pixel* m_pixels;      // page-locked when the device runtime grants it: the frame lands here by one DMA per render
    int m_width, m_height;
    bool m_pinned;
};
#endif
