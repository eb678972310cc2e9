// openskystacker-cl without Qt: same flags, same stdout as cl/cl.cpp:63-225.
//   -f <list.json>  -s  -o <output>  -t <threshold 1-100, default 20>  -j <threads, default 1>  -v
// Extra (not in the reference): --device N, --devices a,b,c | all (the lights are sharded over these GPUs, light k -> GPU k % N like
// the reference's -j workers, util.cpp:738; masters in row slices, one NCCL reduce), --balance (lights from one shared queue), --mono16 (write a single-channel
// result as 16-bit like the GUI does, openskystacker/ui/mainwindow.cpp:91-101).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "imagestacker.hpp"

using namespace openskystacker;

static void usage(int code)
{
    printf("Usage: openskystacker-cl [options]\nMulti-platform deep-sky stacker for astrophotography.\n\nOptions:\n"
           "  -h, --help      Displays this help.\n"
           "  -f <list>       Image list JSON file.\n"
           "  -s              Detect and print the number of stars in the reference\n"
           "                  image with the given threshold, then exit. Ignores all\n"
           "                  other options except -f and -t.\n"
           "  -o <output>     Output image file.\n"
           "  -t <threshold>  Star detection threshold (1-100). Default: 20\n"
           "  -j <threads>    Number of processing threads. Default: 1\n"
           "  -v, --verbose   Verbose output.\n"
           "  --device <n>    CUDA device to run on. Default: 0\n"
           "  --devices <l>   Comma-separated CUDA devices (or 'all') to shard the\n"
           "                  light frames over.\n"
           "  --balance       With several devices: lights from one shared queue\n"
           "                  instead of light k -> device k mod N (unequal host links).\n"
           "  --mono16        Write a single-channel result as a 16-bit TIFF, like\n"
           "                  the GUI does.\n");
    exit(code);
}

int main(int argc, char **argv)
{
    std::string listFile, output, thresholdArg = "20", threadsArg = "1";
    bool detect = false, verbose = false, haveList = false, haveOutput = false, mono16 = false, balance = false;
    int device = 0;
    std::string devicesArg;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto value = [&]() -> std::string { if (i + 1 >= argc) usage(1); return argv[++i]; };
        if (a == "-h" || a == "--help") usage(0);
        else if (a == "-f") { listFile = value(); haveList = true; }
        else if (a == "-s") detect = true;
        else if (a == "-o") { output = value(); haveOutput = true; }
        else if (a == "-t") thresholdArg = value();
        else if (a == "-j") threadsArg = value();
        else if (a == "-v" || a == "--verbose") verbose = true;
        else if (a == "--device") device = atoi(value().c_str());
        else if (a == "--devices") devicesArg = value();
        else if (a == "--mono16") mono16 = true;
        else if (a == "--balance") balance = true;
        else { printf("Unknown option '%s'.\n", a.c_str()); usage(1); }
    }
    if (!haveList) { printf("Error: Must specify an image list\n\n"); usage(1); }
    char *end = nullptr;
    long threshold = strtol(thresholdArg.c_str(), &end, 10);
    if (*end || thresholdArg.empty() || threshold < 1 || threshold > 100) {
        printf("Error: Thread count argument must be an integer between 1 and 100\n"); // sic, cl/cl.cpp:100
        usage(1);
    }
    int err = 0;
    std::vector<ImageRecord *> records = loadImageList(listFile, &err);
    if (err) { printf("Error %d: Couldn't load image list\n", err); return 1; }

    std::vector<int> devices;
    if (devicesArg == "all") {
        int n = visibleDeviceCount();
        for (int i = 0; i < n; i++) devices.push_back(i);
    } else if (!devicesArg.empty()) {
        size_t pos = 0;
        while (pos <= devicesArg.size()) {
            size_t comma = devicesArg.find(',', pos);
            std::string tok = devicesArg.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
            char *e2 = nullptr;
            long v = strtol(tok.c_str(), &e2, 10);
            if (tok.empty() || *e2 || v < 0) { printf("Error: --devices takes a comma-separated list of device numbers or 'all'\n"); usage(1); }
            devices.push_back((int)v);
            if (comma == std::string::npos) break;
            pos = comma + 1;
        }
    }
    ImageStacker stacker(devices.empty() ? device : devices[0]);
    if (!devices.empty()) stacker.setDevices(devices);
    stacker.setDynamicSharding(balance);
    std::string ref;
    std::vector<std::string> lights, darks, darkFlats, flats, bias;
    bool referenceSet = false;
    for (ImageRecord *r : records) if (r->reference) referenceSet = true;
    for (ImageRecord *r : records) { // cl/cl.cpp:126-164
        if (!r->checked) continue;
        switch (r->type) {
        case ImageRecord::LIGHT:
            if (!referenceSet || r->reference) { ref = r->filename; referenceSet = true; }
            else lights.push_back(r->filename);
            break;
        case ImageRecord::DARK: darks.push_back(r->filename); stacker.setUseDarks(true); break;
        case ImageRecord::DARK_FLAT: darkFlats.push_back(r->filename); stacker.setUseDarkFlats(true); break;
        case ImageRecord::FLAT: flats.push_back(r->filename); stacker.setUseFlats(true); break;
        case ImageRecord::BIAS: bias.push_back(r->filename); stacker.setUseBias(true); break;
        }
    }
    for (ImageRecord *r : records) delete r;

    int rc = 0;
    size_t maxMessageLength = 0;
    const int progressBarWidth = 30;
    auto printProgressBar = [&](const std::string &message, int percentage) { // cl/cl.cpp:33-57
        if (verbose) { printf("%d%% %s\n", percentage, message.c_str()); return; }
        char head[16];
        snprintf(head, sizeof head, " %3d%% [", percentage);
        std::string p = head;
        int pos = (int)(percentage / 100.0f * progressBarWidth);
        for (int i = 0; i < progressBarWidth; i++) p += i < pos ? "=" : (i == pos ? ">" : " ");
        p += "] " + message;
        if (p.size() > maxMessageLength) maxMessageLength = p.size();
        p.resize(maxMessageLength, ' ');
        std::cout << p << "\r" << std::flush;
    };
    stacker.updateProgress = printProgressBar;
    stacker.processingError = [&](const std::string &m) { printf("\nError: %s\n", m.c_str()); rc = 1; };

    if (detect) { // cl/cl.cpp:166-169,215-219
        stacker.doneDetectingStars = [](int stars) { printf("%d\n", stars); };
        stacker.detectStars(ref, (int)threshold);
        return rc;
    }
    if (!haveOutput) { printf("Error: Must specify an output image\n"); usage(1); }
    if (output.size() < 4 || output.compare(output.size() - 4, 4, ".tif")) output += ".tif";
    long threads = strtol(threadsArg.c_str(), &end, 10);
    if (*end || threadsArg.empty() || threads < 1) {
        printf("Error: Thread count argument must be a positive non-zero integer\n");
        usage(1);
    }
    stacker.setRefImageFileName(ref);
    stacker.setTargetImageFileNames(lights);
    stacker.setDarkFrameFileNames(darks);
    stacker.setDarkFlatFrameFileNames(darkFlats);
    stacker.setFlatFrameFileNames(flats);
    stacker.setBiasFrameFileNames(bias);
    stacker.finished = [&](const Image &image, const std::string &message) { // cl/cl.cpp:200-213
        printProgressBar(message, 100);
        printf("\n");
        std::string why;
        const bool ok = mono16 ? saveImageLikeGui(output, image, &why)
                               : writeTiffF32(output, image.data.data(), image.width, image.height, image.channels, &why);
        if (!ok) {
            printf("Error: Couldn't write resulting image\n");
            rc = 1;
            return;
        }
        printf("File saved to %s\n", output.c_str());
    };
    stacker.process((int)threshold, (int)threads);
    return rc;
}
