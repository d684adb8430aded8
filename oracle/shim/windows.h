// typedef-only stand-in for <windows.h>: reference include/camera.hpp:4 includes window.hpp,
// which only needs these names to parse. Nothing Win32 is compiled or linked.
#pragma once
#include <cstdint>
typedef void* HWND; typedef void* HDC; typedef void* HBITMAP; typedef void* HINSTANCE; typedef void* HGDIOBJ;
typedef unsigned int UINT; typedef long LONG; typedef unsigned long DWORD; typedef unsigned short WORD;
typedef intptr_t LRESULT; typedef uintptr_t WPARAM; typedef intptr_t LPARAM; typedef int BOOL;
typedef const char* LPCSTR; typedef const wchar_t* LPCWSTR;
#define CALLBACK
#define WINAPI
struct BITMAPINFOHEADER { DWORD biSize; LONG biWidth, biHeight; WORD biPlanes, biBitCount; DWORD biCompression, biSizeImage; LONG biXPelsPerMeter, biYPelsPerMeter; DWORD biClrUsed, biClrImportant; };
struct RGBQUAD { unsigned char b, g, r, x; };
struct BITMAPINFO { BITMAPINFOHEADER bmiHeader; RGBQUAD bmiColors[1]; };
