// Declarations of the reference's assignment-2 scene scripts (the reference's assignment2.h holds exactly these five
// prototypes).  Test infrastructure: lets the reference's UNMODIFIED assignment2.cpp be compiled against the host API
// layer (cse168-raytracer_b200/csrc/miro/) without any other reference header on the include path.
void makeTeapotScene();
void makeBunny1Scene();
void makeBunny20Scene();
void makeSponzaScene();
void makeCornellScene();
