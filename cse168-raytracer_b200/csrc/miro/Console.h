// Console.h -- the reference's logging calls (Console.h:3-6, Console.cpp:67-129): printf-style, fatal() exits with -1.
#ifndef MIROHOST_CONSOLE_H
#define MIROHOST_CONSOLE_H
void warning(const char* fmt, ...);
void error(const char* fmt, ...);
void debug(const char* fmt, ...);
void fatal(const char* fmt, ...);
#endif
