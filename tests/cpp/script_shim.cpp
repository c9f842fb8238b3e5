// Test-only entry into the reference's scene scripts once they are compiled against the host API layer: script_make(name)
// runs make*Scene(), which fills the layer's g_scene / g_camera / g_image; the mh_* functions of libmiro_host.so then
// drive that scene (render, dump, trace).  Returns 0, -1 for an unknown name.
#include <cstring>
#include "assignment2.h"
extern "C" int script_make(const char* name)
{
    if (!strcmp(name, "teapot")) makeTeapotScene();
    else if (!strcmp(name, "bunny1")) makeBunny1Scene();
    else if (!strcmp(name, "bunny20")) makeBunny20Scene();
    else if (!strcmp(name, "cornell")) makeCornellScene();
    else return -1;
    return 0;
}
