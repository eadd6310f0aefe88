from frisk_b200.api import main

if __name__ == "__main__":
    main()
