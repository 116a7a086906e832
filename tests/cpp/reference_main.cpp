// reference_main.cpp -- TEST PROGRAM. What the reference's main.cpp / CLEngineBase::renderLoop do
// (main.cpp:5-13, CLEngineBase.cpp:166-211), written against the host mirror's header exactly as a program
// written against the reference's classes would be: global `eng`, CLEngineBase::renderLoop, CLRaytracer::pixels.
// Usage: reference_main scene.obj width height frames bounces out.raw [device,device,...]
// With a device list the one CLContext drives all of them (multi-GPU frame, no Python anywhere).
#include <cstdio>
#include <cstdlib>
#include <exception>
#include "glaze3d.h"

using namespace Glaze3D;

int main(int argc, char** argv)
{
    if (argc < 7) { std::fprintf(stderr, "usage: %s scene.obj width height frames bounces out.raw\n", argv[0]); return 2; }
    try
    {
        eng = std::make_shared<CLEngineBase>();
        eng->ui->window_width = std::atoi(argv[2]);
        eng->ui->window_height = std::atoi(argv[3]);
        eng->render->lightBounces = std::atoi(argv[5]);
        if (argc > 7)
            for (const char* p = argv[7]; *p;) { eng->render->devices.push_back((int)std::strtol(p, const_cast<char**>(&p), 10)); if (*p == ',') ++p; }
        eng->renderLoop(argv[1], (unsigned)std::atoi(argv[4]));             // Init, Load, CreateBVHTrees, RenderFrame x frames
        const std::vector<float3>& px = eng->render->pixels;
        FILE* f = std::fopen(argv[6], "wb");
        if (!f || std::fwrite(px.data(), sizeof(float3), px.size(), f) != px.size()) { std::fprintf(stderr, "cannot write %s\n", argv[6]); return 3; }
        std::fclose(f);
        std::printf("frames %u, last FPS %.1f, %zu pixels, %d device(s)\n", eng->render->m_FrameCount - 1, eng->FPS, px.size(),
                    b2rt_group_size(eng->render->m_CLContext->GetContext()));
        eng.reset();
    }
    catch (const std::exception& e)
    {
        // the reference shows this text in a MessageBox (CLEngineBase.cpp:181-185)
        std::fprintf(stderr, "CLException: %s\n", e.what());
        return 1;
    }
    return 0;
}
