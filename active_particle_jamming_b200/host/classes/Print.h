// Output files of a run: the same tree, file names and line formats as the reference's Print
// (reference code/classes/Print.h:41-214) so that code/postprocessing/postprocessing.py and the
// gnuplot helpers read the results unchanged:
//
//   <location>{local,remote}_output/<fullRun>/<run>/dat/{COM,orientation,order,corr,orientationCorr,
//       pairCorr,velDist,autoCorr,fluct,densDist,MSD,summary,summary2}.dat   and   .../vid/ovito.txt
//
// Numbers go through operator<< with the stream defaults (6 significant digits), columns are
// tab separated. Directories are made with mkdir(2) instead of system("mkdir"); the post-processing
// scripts are copied next to the results when <location>code/postprocessing exists, as the
// reference does on first use of a <fullRun> directory (Print.h:52-60).
#ifndef APJ_HOST_PRINT_H
#define APJ_HOST_PRINT_H

#include <iostream>
#include <fstream>
#include <string>
#include <iomanip>
#include <sys/stat.h>
#include <sys/types.h>
#include <dirent.h>
#include <cerrno>

struct Print
{
    Print(string, string, string, int, bool);
    ~Print();

    void print_COM(long int, vector<double>&);
    void print_orientation(long int, vector<double>&);
    void print_order(long int, double);
    void print_velDist(double, double);
    void print_corr(double, double);
    void print_orientationCorr(double, double);
    void print_dens(double, double);
    void print_pairCorr(double, double);
    void print_autoCorr(int, double);
    void print_MSD(int, double);
    void print_fluct(double, double, double);
    void print_Ovito(int&, int&, int&, double&, int&, vector<double>&, vector<double>&);
    void print_summary(string, int, double, long int, int, double, double, double, double, long int, double, double, double);

    ofstream COM, orientation, order,
             corr, orientationCorr, pairCorr, autoCorr,
             velDist, MSD, fluct, dens,
             OvitoVid, summary, summary2;

    string run;
    string path;

private:
    static bool make_dir(const string& d) { return mkdir(d.c_str(), 0777) == 0 || errno == EEXIST; }
    static void copy_tree_flat(const string& from, const string& to);
};

inline void Print::copy_tree_flat(const string& from, const string& to)
{
    DIR* d = opendir(from.c_str());
    if (!d) return;
    while (dirent* ent = readdir(d)) {
        const string name = ent->d_name;
        if (name == "." || name == "..") continue;
        ifstream in((from + "/" + name).c_str(), ios::binary);
        if (!in) continue;
        ofstream out((to + name).c_str(), ios::binary);
        out << in.rdbuf();
    }
    closedir(d);
}

inline Print::Print(string location, string fullRun, string ID, int noCells, bool remote)
{
    (void)noCells;
    run = ID;
    const string root = location + (remote ? "remote_output/" : "local_output/");
    path = root + fullRun + "/";

    DIR* existing = opendir(path.c_str());
    if (existing) {
        closedir(existing);
    } else {
        make_dir(location);
        make_dir(root);
        if (!make_dir(path)) {
            cout << "Failed to create folder for the full run. Status 715\n";
            exit(715);
        }
        copy_tree_flat(location + "code/postprocessing", path);
    }

    const string base = path + run;
    bool ok = make_dir(base);
    ok = make_dir(base + "/dat") && ok;
    ok = make_dir(base + "/eps") && ok;
    ok = make_dir(base + "/vid") && ok;
    if (!ok) {
        cout << "Directories not successfully created. Status 716\n";
        exit(716);
    }

    struct { ofstream* f; const char* name; } files[] = {
        {&COM, "/dat/COM.dat"}, {&orientation, "/dat/orientation.dat"}, {&order, "/dat/order.dat"},
        {&corr, "/dat/corr.dat"}, {&orientationCorr, "/dat/orientationCorr.dat"}, {&pairCorr, "/dat/pairCorr.dat"},
        {&velDist, "/dat/velDist.dat"}, {&autoCorr, "/dat/autoCorr.dat"}, {&fluct, "/dat/fluct.dat"},
        {&dens, "/dat/densDist.dat"}, {&MSD, "/dat/MSD.dat"}, {&OvitoVid, "/vid/ovito.txt"},
        {&summary, "/dat/summary.dat"}, {&summary2, "/dat/summary2.dat"}};
    for (auto& e : files) e.f->open((base + e.name).c_str());
}

inline Print::~Print()
{
    ofstream* all[] = {&COM, &orientation, &order, &corr, &orientationCorr, &pairCorr, &autoCorr,
                       &velDist, &fluct, &dens, &MSD, &OvitoVid, &summary, &summary2};
    for (ofstream* f : all) f->close();
}

// ---- time series, one line per sample (reference Print.h:112-127, :166-176) ----
inline void Print::print_COM(long int t, vector<double>& center)
{
    COM << t << "\t" << center[0] << "\t" << center[1] << endl;
}

inline void Print::print_orientation(long int t, vector<double>& orient)
{
    orientation << t << "\t" << orient[0] << "\t" << orient[1] << endl;
}

inline void Print::print_order(long int t, double o)
{
    order << t << "\t" << o << endl;
}

inline void Print::print_MSD(int t, double msd)
{
    MSD << t << "\t" << msd << "\t" << log(msd) << "\t" << log(1.0 - msd) << endl;
}

inline void Print::print_fluct(double r, double avg, double f)
{
    (void)r;    // the reference writes only the expected area and the rms fluctuation (Print.h:174-176)
    fluct << avg << "\t" << f << endl;
}

inline void Print::print_autoCorr(int t, double vaf)
{
    autoCorr << t << "\t" << vaf << endl;
}

// ---- end-of-run tables (reference Print.h:150-164, :178-180) ----
inline void Print::print_corr(double r, double v) { corr << r << "\t" << v << endl; }
inline void Print::print_orientationCorr(double r, double v) { orientationCorr << r << "\t" << v << endl; }
inline void Print::print_pairCorr(double r, double gr) { pairCorr << r << "\t" << gr << "\n"; }
inline void Print::print_velDist(double v, double prob) { velDist << v << "\t" << prob << "\n"; }
inline void Print::print_dens(double density, double count) { dens << density << "\t" << count << endl; }

// ---- Ovito "XYZ" frames: index R over x y vx vy (reference Print.h:129-148) ----
inline void Print::print_Ovito(int& k, int& noCells, int& cellIndex, double& radius, int& overlap,
                               vector<double>& x, vector<double>& v)
{
    if (k == 0) {
        OvitoVid << noCells << endl;
        OvitoVid << "time step comment" << endl;
    }
    OvitoVid << cellIndex << "\t" << radius << "\t" << overlap << "\t" << x[0] << "\t"
             << x[1] << "\t" << v[0] << "\t" << v[1] << endl;
}

// ---- summary.dat (labelled) and summary2.dat (values only), reference Print.h:182-214 ----
inline void Print::print_summary(string ID, int noCells, double L, long int numberOfSteps,
                                 int stepsPerTime, double C1, double C2, double rho,
                                 double seconds, long int resetCounter,
                                 double binder, double order, double variance)
{
    static const char* label[13] = {
        "Run ID:                     ", "Number of cells:            ", "Grid length:                ",
        "Number of steps:            ", "Steps per unit time:        ", "lambda_s (self-propulsion): ",
        "lambda_n (noise):           ", "Rho (density):              ", "Simulation time (seconds):  ",
        "Number of list refreshes:   ", "Binder cumulant:            ", "Average order parameter:    ",
        "Order parameter variance:   "};
    int row = 0;
    auto both = [&](auto value) {
        summary << label[row++] << "\t" << value << endl;
        summary2 << value << endl;
    };
    both(ID); both(noCells); both(L); both(numberOfSteps); both(stepsPerTime); both(C1); both(C2); both(rho);
    both(seconds); both(resetCounter); both(binder); both(order); both(variance);
}

#endif
